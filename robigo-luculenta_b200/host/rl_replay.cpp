// rl_replay.cpp -- drives the units the way the reference's host does.
//
// The reference's host (out of scope for this engine, kept unchanged by the
// integration) is C worker threads that each loop
//     task = scheduler.lock().get_new_task(task); execute_task(task)
// (app.rs:95-126) over a scheduler that recycles 3C trace units and
// max(1, C/2) plot units (task_scheduler.rs:91-182).  This harness replays that
// call pattern against the C++ unit mirrors (rl_units.hpp), from real threads,
// so the threading contract of the C ABI -- distinct handles used concurrently,
// one handle never shared -- is exercised exactly as the Rust host would.
//
//   rl_replay --width W --height H --threads C --batches B [--batch N] [--seed S]
//             [--mode strict|device] [--scene 1..4] [--out PREFIX] [--pin 0|1] [--lazy 0|1] [--async-render 0|1]
//             [--records host|deferred] [--consume device|host]
//             [--first-batch K] [--sleep-ms 100]
// --first-batch: photon ids start at K * batch (a second process -- another GPU -- continues
// the id range of the first).
//
// strict: every call goes through host buffers exactly as app.rs:132-164 does
//         (mapped_photons Vec -> plot(&[MappedPhoton]) -> tristimulus_buffer Vec
//         -> accumulate(&[Vector3]) -> tonemap(&[Vector3])).
// device: the same schedule with the device-resident overloads (no host trips).
// Stops after B trace batches, gathers what is left, tonemaps once, writes
// PREFIX.raw (buffer.raw format) and PREFIX.ppm, prints one JSON line.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rl_host.h"
#include "rl_units.hpp"

using namespace robigo;

namespace {

enum class Kind { Sleep, Trace, Plot, Gather, Tonemap };

struct Task {
    Kind kind = Kind::Sleep;
    std::unique_ptr<TraceUnit> trace;
    std::unique_ptr<PlotUnit> plot;
    std::vector<std::unique_ptr<TraceUnit>> traces;
    std::vector<std::unique_ptr<PlotUnit>> plots;
    std::unique_ptr<GatherUnit> gather;
    std::unique_ptr<TonemapUnit> tonemap;
};

// The unit-recycling policy of task_scheduler.rs:127-182, with the 30 s tonemap
// timer replaced by "when the trace budget is spent" (the harness must end).
class Scheduler {
public:
    Scheduler(size_t concurrency, uint32_t w, uint32_t h, uint64_t seed, uint64_t batch, bool device,
              uint64_t budget, const std::string &raw_path)
        : n_trace_units_(concurrency * 3), budget_(budget) {
        for (size_t i = 0; i < n_trace_units_; i++)
            available_trace_.push_back(std::make_unique<TraceUnit>(i, w, h, seed, batch, device));
        size_t n_plot = concurrency / 2 > 1 ? concurrency / 2 : 1;
        for (size_t i = 0; i < n_plot; i++) available_plot_.push_back(std::make_unique<PlotUnit>(i, w, h, !device));
        raw_path_ = raw_path;
        gather_ = std::make_unique<GatherUnit>(w, h, raw_path_.c_str(), !device);
        tonemap_ = std::make_unique<TonemapUnit>(w, h);
    }

    bool finished() const { return finished_; }
    uint64_t traces_completed() const { return traces_completed_; }
    std::unique_ptr<TonemapUnit> &tonemap_unit() { return tonemap_; }
    std::unique_ptr<GatherUnit> &gather_unit() { return gather_; }
    uint64_t rays() {
        uint64_t total = 0;
        for (auto &u : available_trace_) total += u->ray_count();
        for (auto &u : done_trace_) total += u->ray_count();
        return total;
    }

    Task get_new_task(Task completed) {
        complete(std::move(completed));
        const bool draining = traces_started_ >= budget_;
        if (draining && in_flight_ == 0 && done_trace_.empty() && done_plot_.empty() && gather_ && tonemap_) {
            if (image_changed_) return make_tonemap();
            finished_ = true;
            return Task{};
        }
        if (!draining) {
            if (done_trace_.size() > n_trace_units_ / 2 && !available_plot_.empty()) return make_plot();   // :153-156
            if (!available_trace_.empty()) return make_trace();                                            // :159-161
        }
        if (!available_plot_.empty() && !done_trace_.empty()) return make_plot();                          // :165-168
        if (gather_ && !done_plot_.empty()) return make_gather();                                          // :173-175
        return Task{};                                                                                     // :179 Sleep
    }

private:
    Task make_trace() {
        Task t; t.kind = Kind::Trace;
        t.trace = std::move(available_trace_.front()); available_trace_.pop_front();
        traces_started_++; in_flight_++;
        return t;
    }
    Task make_plot() {                                      // task_scheduler.rs:192-207
        Task t; t.kind = Kind::Plot;
        t.plot = std::move(available_plot_.front()); available_plot_.pop_front();
        size_t n = done_trace_.size() / 2 > 1 ? done_trace_.size() / 2 : 1;
        for (size_t i = 0; i < n && !done_trace_.empty(); i++) {
            t.traces.push_back(std::move(done_trace_.front())); done_trace_.pop_front();
        }
        in_flight_++;
        return t;
    }
    Task make_gather() {                                    // task_scheduler.rs:209-219
        Task t; t.kind = Kind::Gather;
        t.gather = std::move(gather_);
        while (!done_plot_.empty()) { t.plots.push_back(std::move(done_plot_.front())); done_plot_.pop_front(); }
        in_flight_++;
        return t;
    }
    Task make_tonemap() {                                   // task_scheduler.rs:221-228
        Task t; t.kind = Kind::Tonemap;
        t.gather = std::move(gather_); t.tonemap = std::move(tonemap_);
        in_flight_++;
        return t;
    }
    void complete(Task t) {                                 // task_scheduler.rs:231-326
        switch (t.kind) {
        case Kind::Sleep: return;
        case Kind::Trace: done_trace_.push_back(std::move(t.trace)); traces_completed_++; break;
        case Kind::Plot:
            for (auto &u : t.traces) available_trace_.push_back(std::move(u));
            done_plot_.push_back(std::move(t.plot));
            break;
        case Kind::Gather:
            for (auto &u : t.plots) available_plot_.push_back(std::move(u));
            gather_ = std::move(t.gather);
            image_changed_ = true;
            break;
        case Kind::Tonemap:
            gather_ = std::move(t.gather); tonemap_ = std::move(t.tonemap);
            image_changed_ = false;
            break;
        }
        in_flight_--;
    }

    size_t n_trace_units_;
    uint64_t budget_, traces_started_ = 0, traces_completed_ = 0;
    int in_flight_ = 0;
    bool image_changed_ = false, finished_ = false;
    std::deque<std::unique_ptr<TraceUnit>> available_trace_, done_trace_;
    std::deque<std::unique_ptr<PlotUnit>> available_plot_, done_plot_;
    std::unique_ptr<GatherUnit> gather_;
    std::unique_ptr<TonemapUnit> tonemap_;
    std::string raw_path_;
};

// wall time the workers spent in each kind of task (summed over threads), for the report
struct KindStats { std::atomic<uint64_t> ns{0}, calls{0}; };
KindStats g_stats[5];
int g_sleep_ms = 100;   // app.rs:129: the idle task sleeps 100 ms

void execute_kind(Task &t, const Scene &scene, bool device);

// app.rs:113-164
void execute(Task &t, const Scene &scene, bool device) {
    const auto t0 = std::chrono::steady_clock::now();
    execute_kind(t, scene, device);
    const auto dt = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0);
    g_stats[(int)t.kind].ns += (uint64_t)dt.count();
    g_stats[(int)t.kind].calls++;
}

void execute_kind(Task &t, const Scene &scene, bool device) {
    switch (t.kind) {
    case Kind::Sleep: std::this_thread::sleep_for(std::chrono::milliseconds(g_sleep_ms)); break;   // app.rs:128-130
    case Kind::Trace: t.trace->render(scene); break;                                      // app.rs:132-134
    case Kind::Plot:                                                                      // app.rs:136-141
        for (auto &u : t.traces) {
            if (device) t.plot->plot(*u); else t.plot->plot(u->mapped_photons);
        }
        break;
    case Kind::Gather:                                                                    // app.rs:143-152
        for (auto &u : t.plots) {
            if (device) t.gather->accumulate(*u, true);
            else { t.gather->accumulate(u->tristimulus_buffer); u->clear(); }
        }
        t.gather->save();
        break;
    case Kind::Tonemap:                                                                   // app.rs:154-164
        if (device) t.tonemap->tonemap(*t.gather); else t.tonemap->tonemap(t.gather->tristimulus_buffer);
        break;
    }
}

const char *arg(int argc, char **argv, const char *name, const char *fallback) {
    for (int i = 1; i + 1 < argc; i++)
        if (!strcmp(argv[i], name)) return argv[i + 1];
    return fallback;
}

}  // namespace

int main(int argc, char **argv) {
    const uint32_t w = (uint32_t)atoi(arg(argc, argv, "--width", "1280"));      // main.rs:47-48
    const uint32_t h = (uint32_t)atoi(arg(argc, argv, "--height", "720"));
    const size_t threads = (size_t)atoi(arg(argc, argv, "--threads", "4"));
    const uint64_t batches = strtoull(arg(argc, argv, "--batches", "16"), nullptr, 10);
    const uint64_t batch = strtoull(arg(argc, argv, "--batch", "524288"), nullptr, 10);
    const uint64_t seed = strtoull(arg(argc, argv, "--seed", "24301"), nullptr, 10);
    const int which = atoi(arg(argc, argv, "--scene", "2"));
    const bool device = !strcmp(arg(argc, argv, "--mode", "strict"), "device");
    const std::string out = arg(argc, argv, "--out", "replay");
    pin_host_buffers() = atoi(arg(argc, argv, "--pin", "1")) != 0;
    lazy_host_mirrors() = atoi(arg(argc, argv, "--lazy", "1")) != 0;
    async_render() = atoi(arg(argc, argv, "--async-render", "1")) != 0;
    deferred_records() = !strcmp(arg(argc, argv, "--records", "host"), "deferred");
    consume_on_device() = strcmp(arg(argc, argv, "--consume", "device"), "host") != 0;
    g_sleep_ms = atoi(arg(argc, argv, "--sleep-ms", "100"));

    try {
        rl_scene_builder *builder = nullptr;
        expect(rl_scene_builder_create(&builder), "rl_scene_builder_create");
        expect(rl_scene_builder_builtin(builder, which, 0), "rl_scene_builder_builtin");
        rl_scene_desc desc;
        expect(rl_scene_builder_desc(builder, &desc), "rl_scene_builder_desc");
        Scene scene(desc);
        {
            // untimed warm-up: loads the kernels (one small batch through every unit type)
            TraceUnit tu(0, 64, 64, seed, 4096, false);
            PlotUnit pu(0, 64, 64);
            GatherUnit gu(64, 64, nullptr);
            TonemapUnit mu(64, 64);
            tu.render(scene);
            pu.plot(tu.mapped_photons);
            gu.accumulate(pu.tristimulus_buffer);
            mu.tonemap(gu.tristimulus_buffer);
        }
        expect(rl_scene_batch_counter_reset(scene.handle(), strtoull(arg(argc, argv, "--first-batch", "0"), nullptr, 10)),
               "rl_scene_batch_counter_reset");
        const std::string raw = out + ".raw";
        remove(raw.c_str());   // start from black: GatherUnit::new resumes from the file if present

        // unit creation (TaskScheduler::new, task_scheduler.rs:91-125: 3C trace units, C/2 plot units,
        // page-locking of their host buffers) is timed separately: once per render, like App::new
        const auto t_setup = std::chrono::steady_clock::now();
        Scheduler scheduler(threads, w, h, seed, batch, device, batches, raw);
        const double setup_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_setup).count();
        std::mutex lock;
        rl_transfer_counters_reset();
        const auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> workers;
        for (size_t i = 0; i < threads; i++) {
            workers.emplace_back([&]() {                                        // app.rs:95-111
                Task task;
                for (;;) {
                    {
                        std::lock_guard<std::mutex> guard(lock);
                        if (scheduler.finished()) return;
                        task = scheduler.get_new_task(std::move(task));
                        if (scheduler.finished()) return;
                    }
                    execute(task, scene, device);
                }
            });
        }
        for (auto &t : workers) t.join();
        // the run ends when buffer.raw is on disk (the writer thread of the gather unit)
        expect(rl_gather_unit_flush(scheduler.gather_unit()->handle()), "rl_gather_unit_flush");
        const double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        uint64_t h2d = 0, d2h = 0;
        rl_transfer_counters(&h2d, &d2h);

        const auto &rgb = scheduler.tonemap_unit()->rgb_buffer;
        FILE *f = fopen((out + ".ppm").c_str(), "wb");
        if (f) {
            fprintf(f, "P6\n%u %u\n255\n", w, h);
            fwrite(rgb.data(), 1, rgb.size(), f);
            fclose(f);
        }
        const uint64_t rays = scheduler.rays();
        uint64_t launches = 0, grouped = 0;
        rl_scene_dispatch_stats(scene.handle(), &launches, &grouped);
        printf("{\"mode\": \"%s\", \"pinned\": %s, \"lazy\": %s, \"async_render\": %s, \"records\": \"%s\", \"consume\": \"%s\", \"sleep_ms\": %d, \"width\": %u, \"height\": %u, \"threads\": %zu, \"batches\": %llu, \"batch\": %llu, \"seconds\": %.6f, \"setup_seconds\": %.6f, "
               "\"batches_per_s\": %.3f, \"dispatch\": {\"launches\": %llu, \"batches\": %llu}, \"rays\": %llu, \"mrays_per_s\": %.3f, \"h2d_bytes\": %llu, \"d2h_bytes\": %llu, \"worker_seconds\": {\"sleep\": [%llu, %.3f], "
               "\"trace\": [%llu, %.3f], \"plot\": [%llu, %.3f], \"gather\": [%llu, %.3f], \"tonemap\": [%llu, %.3f]}}\n",
               device ? "device" : "strict", pin_host_buffers() ? "true" : "false",
               lazy_host_mirrors() ? "true" : "false", async_render() ? "true" : "false",
               deferred_records() ? "deferred" : "host", consume_on_device() ? "device" : "host", g_sleep_ms, w, h, threads, (unsigned long long)scheduler.traces_completed(),
               (unsigned long long)batch, seconds, setup_seconds, scheduler.traces_completed() / seconds,
               (unsigned long long)launches, (unsigned long long)grouped, (unsigned long long)rays, rays / seconds / 1e6, (unsigned long long)h2d, (unsigned long long)d2h,
               (unsigned long long)g_stats[0].calls, g_stats[0].ns * 1e-9, (unsigned long long)g_stats[1].calls,
               g_stats[1].ns * 1e-9, (unsigned long long)g_stats[2].calls, g_stats[2].ns * 1e-9,
               (unsigned long long)g_stats[3].calls, g_stats[3].ns * 1e-9, (unsigned long long)g_stats[4].calls,
               g_stats[4].ns * 1e-9);
        rl_scene_builder_destroy(builder);
    } catch (const std::exception &e) {
        fprintf(stderr, "rl_replay: %s\n", e.what());
        return 1;
    }
    return 0;
}
