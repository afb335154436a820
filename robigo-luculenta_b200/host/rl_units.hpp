// rl_units.hpp -- C++ mirror of the reference's four pipeline units over the
// C ABI (include/rl_b200.h).  Same type names, methods and public fields as the
// Rust structs the host drives (trace_unit.rs:40-168, plot_unit.rs:23-103,
// gather_unit.rs:24-94, tonemap_unit.rs:22-101), so a scheduler written against
// the reference's API compiles against these unchanged.  This is what the Rust
// shims of INTEGRATION.md do, written in the language this image can build.
//
// Failures of the C ABI throw (the reference panics: app.rs:107,163).
#pragma once

#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/rl_b200.h"

namespace robigo {

struct Vector3 { float x, y, z; };                     // vector3.rs:20-25
using MappedPhoton = rl_mapped_photon;                 // trace_unit.rs:23-37

inline void expect(int status, const char *what) {
    if (status != RL_OK) throw std::runtime_error(std::string(what) + ": " + rl_last_error());
}

// The units' host buffers are allocated once and never resized (trace_unit.rs:75-77,
// plot_unit.rs:47, gather_unit.rs:40, tonemap_unit.rs:47): page-lock them once so that the
// copies behind render()/plot()/accumulate() are DMA transfers (rl_host_register).
// `pin_host_buffers() = false` keeps them pageable (A/B measurements).
inline bool &pin_host_buffers() { static bool on = true; return on; }
template <typename T> inline void pin(std::vector<T> &v) {
    if (pin_host_buffers() && !v.empty()) expect(rl_host_register(v.data(), v.size() * sizeof(T)), "rl_host_register");
}
template <typename T> inline void unpin(std::vector<T> &v) {
    if (pin_host_buffers() && !v.empty()) rl_host_unregister(v.data());
}

// A unit's read-only host buffer whose truth lives on the device: `tristimulus_buffer` of
// PlotUnit (plot_unit.rs:34) and GatherUnit (gather_unit.rs:26).  The host only ever reads
// these fields, as slices handed to the next unit (app.rs:146,157), so the copy back from the
// device is made when the field is read, not after every plot()/accumulate() that changes it:
// a plot task plots a dozen trace units into one plot unit (task_scheduler.rs:192-207) before
// anybody looks at the result.  Reads go through the conversions below, which is where a Rust
// shim puts `impl Deref<Target = [Vector3]>` -- `&unit.tristimulus_buffer` at the call sites
// of app.rs coerces to `&[Vector3]` unchanged.  `mapped_photons` of TraceUnit (trace_unit.rs:56)
// is the same kind of field: render() queues the kernel and the copy into it, the first read
// (app.rs:139, in a later plot task) waits for them.  `lazy_host_mirrors() = false` copies back
// eagerly after every change instead (A/B measurements).
inline bool &lazy_host_mirrors() { static bool on = true; return on; }
// TraceUnit::render returns once the batch is queued (true) or once `mapped_photons` is filled
inline bool &async_render() { static bool on = true; return on; }
// `mapped_photons` is copied out of the device by render() (false: the records cross PCIe twice,
// as a literal reading of app.rs:133-139 has it) or only when host code reads it (true:
// PlotUnit::plot recognises the field by its type and splats from the device records)
inline bool &deferred_records() { static bool on = false; return on; }
// A unit's buffer handed to the next unit (`plot(&unit.mapped_photons)`, app.rs:139;
// `accumulate(&plot_unit.tristimulus_buffer)`, app.rs:146; `tonemap(&gather.tristimulus_buffer)`,
// app.rs:157) is consumed where it was produced, on the device (true): the host copy is still
// filled for whoever reads the field, but nothing is uploaded again.  false: the literal
// reading -- the host slice is copied back to the device (A/B measurements).
inline bool &consume_on_device() { static bool on = true; return on; }
template <typename T> class HostMirror {
public:
    using Download = std::function<void(T *)>;
    HostMirror() = default;
    ~HostMirror() { unpin(storage_); }
    HostMirror(const HostMirror &) = delete;
    HostMirror &operator=(const HostMirror &) = delete;
    void init(size_t n, Download download) {
        storage_.assign(n, T{});
        pin(storage_);
        download_ = std::move(download);
    }
    // the device copy changed
    void invalidate() {
        if (!download_) return;
        stale_ = true;
        if (!lazy_host_mirrors()) sync();
    }
    const std::vector<T> &get() const { sync(); return storage_; }
    T *destination() { return storage_.data(); }       // where the device copy lands (no wait)
    // the unit whose device buffer this field mirrors, while that buffer is newer than the host copy
    void set_owner(void *unit_handle) { owner_ = unit_handle; }
    // (the host can only read these fields, so the device copy is what the field holds)
    void *device_owner() const { return consume_on_device() || (stale_ && deferred_records()) ? owner_ : nullptr; }
    operator const std::vector<T> &() const { return get(); }
    const T *data() const { return get().data(); }
    size_t size() const { return storage_.size(); }
    bool empty() const { return storage_.empty(); }
    const T &operator[](size_t i) const { return get()[i]; }
    typename std::vector<T>::const_iterator begin() const { return get().begin(); }
    typename std::vector<T>::const_iterator end() const { return get().end(); }

private:
    void sync() const {
        if (stale_) {
            download_(storage_.data());
            stale_ = false;
        }
    }
    mutable std::vector<T> storage_;
    mutable bool stale_ = false;
    Download download_;
    void *owner_ = nullptr;
};

// Stands in for Arc<Scene> (app.rs:63): the flattened scene on the device.
class Scene {
public:
    explicit Scene(const rl_scene_desc &desc) { expect(rl_scene_create(&desc, &handle_), "rl_scene_create"); }
    ~Scene() { rl_scene_destroy(handle_); }
    Scene(const Scene &) = delete;
    Scene &operator=(const Scene &) = delete;
    const rl_scene *handle() const { return handle_; }

private:
    rl_scene *handle_ = nullptr;
};

class TraceUnit {
public:
    // trace_unit.rs:64-78.  `keep_on_device`: leave the records on the GPU for
    // PlotUnit::plot(TraceUnit&) instead of filling `mapped_photons`.
    TraceUnit(size_t id_, uint32_t width, uint32_t height, uint64_t seed = 0x5EED,
              uint64_t batch = RL_BATCH_PHOTONS, bool keep_on_device = false)
        : id(id_), keep_on_device_(keep_on_device) {
        expect(rl_trace_unit_create(id_, width, height, seed, &handle_), "rl_trace_unit_create");
        expect(rl_trace_unit_set_batch_size(handle_, batch), "rl_trace_unit_set_batch_size");
        if (!keep_on_device) {
            mapped_photons.init(batch, [this](MappedPhoton *dst) {
                if (deferred_records())
                    expect(rl_trace_unit_download(handle_, dst, mapped_photons.size(), nullptr), "rl_trace_unit_download");
                else expect(rl_trace_unit_sync(handle_), "rl_trace_unit_sync");
            });
            mapped_photons.set_owner(handle_);
        }
    }
    ~TraceUnit() { rl_trace_unit_destroy(handle_); }
    TraceUnit(const TraceUnit &) = delete;
    TraceUnit &operator=(const TraceUnit &) = delete;

    // trace_unit.rs:151-168.  The batch is queued (kernel, then the copy into the page-locked
    // `mapped_photons`) and the call returns; whoever reads `mapped_photons` first waits for it
    // (app.rs:139 in the plot task), so the worker thread is free for its next task meanwhile.
    void render(const Scene &scene) {
        expect(rl_trace_unit_render_async(handle_, scene.handle(),
                                          keep_on_device_ || deferred_records() ? nullptr : mapped_photons.destination()),
               "rl_trace_unit_render");
        mapped_photons.invalidate();
        if (!async_render()) {
            if (keep_on_device_ || deferred_records()) expect(rl_trace_unit_sync(handle_), "rl_trace_unit_sync");
            else mapped_photons.get();
        }
    }
    uint64_t ray_count() {
        uint64_t n = 0;
        expect(rl_trace_unit_ray_count(handle_, &n), "rl_trace_unit_ray_count");
        return n;
    }
    rl_trace_unit *handle() { return handle_; }

    HostMirror<MappedPhoton> mapped_photons;           // trace_unit.rs:56
    size_t id;                                         // trace_unit.rs:59

private:
    rl_trace_unit *handle_ = nullptr;
    bool keep_on_device_;
};

class PlotUnit {
public:
    // plot_unit.rs:43-53.  `mirror_on_host`: keep `tristimulus_buffer` readable on the host,
    // as the unchanged app.rs:146 reads the field.
    PlotUnit(size_t id_, uint32_t width, uint32_t height, bool mirror_on_host = true) : id(id_) {
        expect(rl_plot_unit_create(id_, width, height, &handle_), "rl_plot_unit_create");
        if (mirror_on_host) {
            tristimulus_buffer.init((size_t)width * height, [this](Vector3 *dst) {
                expect(rl_plot_unit_download(handle_, reinterpret_cast<float *>(dst)), "rl_plot_unit_download");
            });
            tristimulus_buffer.set_owner(handle_);
        }
    }
    ~PlotUnit() { rl_plot_unit_destroy(handle_); }
    PlotUnit(const PlotUnit &) = delete;
    PlotUnit &operator=(const PlotUnit &) = delete;

    // plot_unit.rs:87-95
    void plot(const std::vector<MappedPhoton> &photons) {
        expect(rl_plot_unit_plot(handle_, photons.data(), photons.size()), "rl_plot_unit_plot");
        tristimulus_buffer.invalidate();
    }
    // `plot(&unit.mapped_photons)` as app.rs:139 writes it: the argument is the trace unit's own
    // field, recognised by its type (Rust: `plot<P: AsPhotons + ?Sized>(&mut self, photons: &P)`,
    // implemented for [MappedPhoton] and for HostMirror<MappedPhoton>).  While the device records
    // are newer than the host copy they are splatted where they are; no byte crosses PCIe.
    void plot(const HostMirror<MappedPhoton> &photons) {
        if (void *owner = photons.device_owner()) {
            expect(rl_plot_unit_plot_device(handle_, static_cast<rl_trace_unit *>(owner)), "rl_plot_unit_plot_device");
            tristimulus_buffer.invalidate();
        } else {
            plot(photons.get());
        }
    }
    // the same on records a trace unit left on the device
    void plot(TraceUnit &unit) {
        expect(rl_plot_unit_plot_device(handle_, unit.handle()), "rl_plot_unit_plot_device");
        tristimulus_buffer.invalidate();
    }
    // plot_unit.rs:98-102
    void clear() {
        expect(rl_plot_unit_clear(handle_), "rl_plot_unit_clear");
        tristimulus_buffer.invalidate();
    }
    rl_plot_unit *handle() { return handle_; }

    HostMirror<Vector3> tristimulus_buffer;            // plot_unit.rs:34
    size_t id;                                         // plot_unit.rs:37

private:
    rl_plot_unit *handle_ = nullptr;
};

class GatherUnit {
public:
    // gather_unit.rs:35-46: resumes from "buffer.raw" if it exists
    GatherUnit(uint32_t width, uint32_t height, const char *resume_path = "buffer.raw", bool mirror_on_host = true)
        : path_(resume_path ? resume_path : "") {
        expect(rl_gather_unit_create(width, height, resume_path, &handle_), "rl_gather_unit_create");
        if (mirror_on_host) {
            tristimulus_buffer.init((size_t)width * height, [this](Vector3 *dst) {
                expect(rl_gather_unit_download(handle_, reinterpret_cast<float *>(dst), nullptr),
                       "rl_gather_unit_download");
            });
            tristimulus_buffer.invalidate();           // the resumed state
            tristimulus_buffer.set_owner(handle_);
        }
    }
    ~GatherUnit() { rl_gather_unit_destroy(handle_); }
    GatherUnit(const GatherUnit &) = delete;
    GatherUnit &operator=(const GatherUnit &) = delete;

    // gather_unit.rs:49-64
    void accumulate(const std::vector<Vector3> &tristimuli) {
        expect(rl_gather_unit_accumulate(handle_, reinterpret_cast<const float *>(tristimuli.data())),
               "rl_gather_unit_accumulate");
        tristimulus_buffer.invalidate();
    }
    // `accumulate(&plot_unit.tristimulus_buffer)` as app.rs:146 writes it: the argument is a plot
    // unit's own field (Rust: `accumulate<T: AsTristimuli + ?Sized>`), whose truth is on the device
    void accumulate(const HostMirror<Vector3> &tristimuli) {
        if (void *owner = tristimuli.device_owner()) {
            expect(rl_gather_unit_accumulate_plot(handle_, static_cast<rl_plot_unit *>(owner), 0),
                   "rl_gather_unit_accumulate_plot");
            tristimulus_buffer.invalidate();
        } else {
            accumulate(tristimuli.get());
        }
    }
    // accumulate(&plot.tristimulus_buffer) + plot.clear() without the host trip (app.rs:145-148)
    void accumulate(PlotUnit &plot, bool clear_plot) {
        expect(rl_gather_unit_accumulate_plot(handle_, plot.handle(), clear_plot ? 1 : 0),
               "rl_gather_unit_accumulate_plot");
        tristimulus_buffer.invalidate();
        if (clear_plot) plot.tristimulus_buffer.invalidate();
    }
    // gather_unit.rs:68-78; the file is written behind the call (rl_gather_unit_save), the
    // destructor waits for it
    void save() {
        if (!path_.empty()) expect(rl_gather_unit_save(handle_, path_.c_str()), "failed to open file");
    }
    rl_gather_unit *handle() { return handle_; }

    HostMirror<Vector3> tristimulus_buffer;            // gather_unit.rs:26

private:
    rl_gather_unit *handle_ = nullptr;
    std::string path_;
};

class TonemapUnit {
public:
    // tonemap_unit.rs:43-51
    TonemapUnit(uint32_t width, uint32_t height) : rgb_buffer((size_t)width * height * 3, 0) {
        expect(rl_tonemap_unit_create(width, height, &handle_), "rl_tonemap_unit_create");
        pin(rgb_buffer);
    }
    ~TonemapUnit() { rl_tonemap_unit_destroy(handle_); unpin(rgb_buffer); }
    TonemapUnit(const TonemapUnit &) = delete;
    TonemapUnit &operator=(const TonemapUnit &) = delete;

    // tonemap_unit.rs:73-100
    void tonemap(const std::vector<Vector3> &tristimuli) {
        expect(rl_tonemap_unit_tonemap(handle_, reinterpret_cast<const float *>(tristimuli.data()), rgb_buffer.data()),
               "rl_tonemap_unit_tonemap");
    }
    // `tonemap(&gather_unit.tristimulus_buffer)` as app.rs:157 writes it
    void tonemap(const HostMirror<Vector3> &tristimuli) {
        if (void *owner = tristimuli.device_owner())
            expect(rl_tonemap_unit_tonemap_gather(handle_, static_cast<rl_gather_unit *>(owner), rgb_buffer.data()),
                   "rl_tonemap_unit_tonemap_gather");
        else
            tonemap(tristimuli.get());
    }
    void tonemap(GatherUnit &gather) {
        expect(rl_tonemap_unit_tonemap_gather(handle_, gather.handle(), rgb_buffer.data()),
               "rl_tonemap_unit_tonemap_gather");
    }

    std::vector<uint8_t> rgb_buffer;                   // tonemap_unit.rs:30

private:
    rl_tonemap_unit *handle_ = nullptr;
};

}  // namespace robigo
