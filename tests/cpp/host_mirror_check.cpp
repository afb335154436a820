// CPU check of the lazy host buffers of the C++ unit mirror (host/rl_units.hpp): HostMirror copies
// back from "the device" exactly when the field is read after a change -- the rule that replaces
// the eager copies behind PlotUnit::plot / GatherUnit::accumulate / TraceUnit::render.  No GPU:
// page-locking is switched off and the download is a counting stub.
#include <cstdio>
#include <cstdlib>

#include "../../robigo-luculenta_b200/host/rl_units.hpp"

using namespace robigo;

#define CHECK(cond)                                                                   \
    do {                                                                              \
        if (!(cond)) { fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } \
    } while (0)

static size_t first_x(const std::vector<Vector3> &v) { return (size_t)v[0].x; }

int main() {
    pin_host_buffers() = false;
    int downloads = 0;
    float device_value = 1.0f;
    HostMirror<Vector3> m;
    CHECK(m.empty() && m.size() == 0);
    m.invalidate();                                  // not initialised: nothing to fetch
    CHECK(downloads == 0);
    m.init(4, [&](Vector3 *dst) { downloads++; for (int i = 0; i < 4; i++) dst[i] = Vector3{device_value, 0.f, 0.f}; });
    CHECK(m.size() == 4 && downloads == 0);
    CHECK(m[0].x == 0.0f && downloads == 0);         // fresh and clean: the zeros of the constructor

    // a dozen changes, one read: one copy (task_scheduler.rs:192-207 plots a dozen units per task)
    for (int i = 0; i < 12; i++) { device_value = 2.0f + i; m.invalidate(); }
    CHECK(downloads == 0);
    CHECK(first_x(m) == 13 && downloads == 1);       // conversion to const std::vector<Vector3>& (app.rs:146)
    CHECK(m.data()[3].x == 13.0f && m.begin()->x == 13.0f && (m.end() - m.begin()) == 4 && downloads == 1);

    // eager mode: every change copies back at once
    lazy_host_mirrors() = false;
    device_value = 99.0f;
    m.invalidate();
    CHECK(downloads == 2 && m[1].x == 99.0f && downloads == 2);
    lazy_host_mirrors() = true;

    // the trace unit's field: the host can only read it, so the device copy is what it holds and
    // plot() is pointed at the owning unit instead of uploading the host copy again; with
    // consume_on_device() off (the literal reading of app.rs:139) only records that were never
    // copied down (deferred records) are taken from the device
    int fake_unit = 0;
    m.set_owner(&fake_unit);
    CHECK(m.device_owner() == &fake_unit);           // default: consumed where it was produced
    consume_on_device() = false;
    CHECK(m.device_owner() == nullptr);              // clean
    m.invalidate();
    CHECK(m.device_owner() == nullptr);              // stale, but records are copied down by render()
    deferred_records() = true;
    CHECK(m.device_owner() == &fake_unit && downloads == 2);
    CHECK(m[0].x == 99.0f && downloads == 3);        // host code reads the field after all: one copy
    CHECK(m.device_owner() == nullptr);              // ... and the host copy is current again
    deferred_records() = false;
    consume_on_device() = true;
    CHECK(m.device_owner() == &fake_unit && downloads == 3);   // no copy to find that out
    puts("host mirror ok");
    return 0;
}
