"""The C++ host mirror (host/rl_units.hpp) driven from real threads by the
scheduler replay (host/rl_replay.cpp): the call pattern of app.rs:95-164 and
task_scheduler.rs:91-182 against the C ABI, checked against the oracle."""
import json
import os
import subprocess

import numpy as np
import pytest

import __graft_entry__ as entry

pytestmark = pytest.mark.gpu


# strict: host buffers as app.rs:132-164 passes them; "deferred": the same unchanged call sites,
# the records leave the device only if host code reads them (PlotUnit::plot recognises the trace
# unit's own field); the extra flags switch the page-locking, lazy mirrors and queued render off
@pytest.mark.parametrize("mode,threads,extra", [
    ("strict", 4, []), ("device", 4, []), ("strict", 1, []), ("device", 7, []),
    ("strict", 5, ["--records", "deferred"]), ("strict", 2, ["--records", "deferred", "--lazy", "0"]),
    ("strict", 3, ["--pin", "0", "--lazy", "0", "--async-render", "0"])])
def test_scheduler_replay_matches_oracle(gpu, orc, tmp_path, mode, threads, extra):
    exe = entry.build_replay()
    w, h, batch, batches, seed = 320, 180, 4096, 24, 24301
    out = str(tmp_path / f"replay_{mode}")
    res = subprocess.run([exe, "--width", str(w), "--height", str(h), "--threads", str(threads), "--batches",
                          str(batches), "--batch", str(batch), "--seed", str(seed), "--mode", mode, "--scene", "2",
                          "--out", out] + extra, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr
    stats = json.loads(res.stdout.strip().splitlines()[-1])
    assert stats["batches"] == batches and stats["mode"] == mode
    assert stats["records"] == ("deferred" if "deferred" in extra else "host")
    # whichever unit rendered which batch, the union of photons is ids [0, batches * batch)
    d = gpu.SceneBuilder(2).desc()
    ct = orc.Counters()
    photons = orc.trace(d, seed, w, h, 0, batches * batch, orc.MATH_SPEC, False, ct)
    assert stats["rays"] == ct.rays
    want = orc.plot(w, h, photons)
    # buffer.raw format: accumulator then compensation, 24*w*h bytes (gather_unit.rs:68-78)
    raw = np.fromfile(out + ".raw", dtype="<f4")
    assert raw.size == 6 * w * h
    acc = raw[: 3 * w * h].reshape(h, w, 3)
    # plot buffers are summed in a schedule-dependent order: compare within float summation noise
    tol = 2e-5 * float(np.abs(want).max()) + 1e-12
    assert float(np.abs(acc - want).max()) <= tol
    ppm = open(out + ".ppm", "rb").read()
    assert ppm.startswith(b"P6\n%d %d\n255\n" % (w, h)) and len(ppm) > 3 * w * h
    rgb = np.frombuffer(ppm[-3 * w * h:], dtype=np.uint8).reshape(h, w, 3)
    ref_rgb = orc.tonemap(np.ascontiguousarray(acc), orc.MATH_LIBM)
    assert np.abs(rgb.astype(int) - ref_rgb.astype(int)).max() <= 1
