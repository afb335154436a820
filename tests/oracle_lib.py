"""ctypes wrapper of oracle/liboracle.so -- the CPU checker (test infrastructure).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
legs import this module; the product never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "oracle", "liboracle.so")

MATH_LIBM, MATH_SPEC = 0, 1

MAPPED_PHOTON = np.dtype([("x", "<f4"), ("y", "<f4"), ("probability", "<f4"), ("wavelength", "<f4")])
RAY = np.dtype([("origin", "<f4", 3), ("direction", "<f4", 3), ("wavelength", "<f4"),
                ("probability", "<f4")])
HIT = np.dtype([("object", "<i4"), ("distance", "<f4"), ("position", "<f4", 3),
                ("normal", "<f4", 3), ("tangent", "<f4", 3)])


class Counters(C.Structure):
    _fields_ = [("photons", C.c_uint64), ("rays", C.c_uint64), ("primitive_tests", C.c_uint64),
                ("emissive_hits", C.c_uint64), ("max_bounces", C.c_uint64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True)
        _lib = C.CDLL(LIB_PATH)
        for name in ("orc_trace", "orc_plot", "orc_gather_accumulate", "orc_find_exposure", "orc_tonemap",
                     "orc_intersect", "orc_math", "orc_blackbody_intensity", "orc_tristimulus",
                     "orc_camera_rays", "orc_philox", "orc_draws", "orc_render_mt", "orc_hardware_threads"):
            getattr(_lib, name).restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _ok(rc):
    if rc != 0:
        raise RuntimeError(f"oracle status {rc}")


def trace(desc, seed, width, height, first, n, mode=MATH_SPEC, count_tests=False, counters=None):
    """TraceUnit::render over photon ids [first, first+n)."""
    out = np.zeros(n, dtype=MAPPED_PHOTON)
    ct = counters if counters is not None else Counters()
    _ok(lib().orc_trace(C.byref(desc), C.c_uint64(seed), C.c_uint32(width), C.c_uint32(height),
                        C.c_uint64(first), C.c_uint64(n), C.c_int(mode), C.c_int(1 if count_tests else 0),
                        _p(out), C.byref(ct)))
    return out


def plot(width, height, photons, xyz=None):
    """PlotUnit::plot into a (h, w, 3) f32 buffer (accumulating)."""
    photons = np.ascontiguousarray(photons, dtype=MAPPED_PHOTON)
    if xyz is None:
        xyz = np.zeros((height, width, 3), dtype=np.float32)
    _ok(lib().orc_plot(C.c_uint32(width), C.c_uint32(height), _p(photons), C.c_uint64(photons.shape[0]), _p(xyz)))
    return xyz


def gather_accumulate(acc, comp, px):
    assert acc.dtype == np.float32 and comp.dtype == np.float32
    px = np.ascontiguousarray(px, dtype=np.float32)
    _ok(lib().orc_gather_accumulate(_p(acc), _p(comp), _p(px), C.c_uint64(acc.size // 3)))


def find_exposure(xyz):
    h, w, _ = xyz.shape
    out = C.c_float()
    x = np.ascontiguousarray(xyz, dtype=np.float32)
    _ok(lib().orc_find_exposure(C.c_uint32(w), C.c_uint32(h), _p(x), C.byref(out)))
    return float(out.value)


def tonemap(xyz, mode=MATH_SPEC, exposure=None):
    h, w, _ = xyz.shape
    x = np.ascontiguousarray(xyz, dtype=np.float32)
    rgb = np.zeros((h, w, 3), dtype=np.uint8)
    e = C.c_float(float("nan") if exposure is None else exposure)
    _ok(lib().orc_tonemap(C.c_uint32(w), C.c_uint32(h), _p(x), C.c_int(mode), e, _p(rgb)))
    return rgb


def intersect(desc, rays):
    rays = np.ascontiguousarray(rays, dtype=RAY)
    out = np.zeros(rays.shape[0], dtype=HIT)
    _ok(lib().orc_intersect(C.byref(desc), _p(rays), C.c_uint64(rays.shape[0]), _p(out)))
    return out


def math(fn, x, x2=None, mode=MATH_SPEC):
    x = np.ascontiguousarray(x, dtype=np.float32)
    x2 = None if x2 is None else np.ascontiguousarray(x2, dtype=np.float32)
    out = np.zeros_like(x)
    _ok(lib().orc_math(C.c_int(fn), C.c_int(mode), _p(x), _p(x2), C.c_uint64(x.size), _p(out)))
    return out


def blackbody_intensity(temperature, normalisation, wavelengths, mode=MATH_SPEC):
    w = np.ascontiguousarray(wavelengths, dtype=np.float32)
    out = np.zeros_like(w)
    _ok(lib().orc_blackbody_intensity(C.c_float(temperature), C.c_float(normalisation), C.c_int(mode),
                                      _p(w), C.c_uint64(w.size), _p(out)))
    return out


def tristimulus(wavelengths):
    w = np.ascontiguousarray(wavelengths, dtype=np.float32)
    out = np.zeros((w.size, 3), dtype=np.float32)
    _ok(lib().orc_tristimulus(_p(w), C.c_uint64(w.size), _p(out)))
    return out


def camera_rays(desc, seed, width, height, first, n, mode=MATH_SPEC):
    rays = np.zeros(n, dtype=RAY)
    xy = np.zeros(n, dtype=MAPPED_PHOTON)
    _ok(lib().orc_camera_rays(C.byref(desc), C.c_uint64(seed), C.c_uint32(width), C.c_uint32(height),
                              C.c_uint64(first), C.c_uint64(n), C.c_int(mode), _p(rays), _p(xy)))
    return rays, xy


def philox(key, ctr):
    out = (C.c_uint32 * 4)()
    _ok(lib().orc_philox(C.c_uint32(key[0]), C.c_uint32(key[1]), C.c_uint32(ctr[0]), C.c_uint32(ctr[1]),
                         C.c_uint32(ctr[2]), C.c_uint32(ctr[3]), out))
    return [int(v) for v in out]


def draws(seed, photon, half_open_flags):
    flags = np.ascontiguousarray(half_open_flags, dtype=np.uint8)
    out = np.zeros(flags.size, dtype=np.float32)
    _ok(lib().orc_draws(C.c_uint64(seed), C.c_uint64(photon), C.c_uint32(flags.size), _p(flags), _p(out)))
    return out


def render_mt(desc, seed, width, height, first, n, threads, mode=MATH_LIBM, batch=0, want_image=True):
    """Multi-threaded trace+plot(+gather): returns (xyz or None, counters dict, seconds)."""
    xyz = np.zeros((height, width, 3), dtype=np.float32) if want_image else None
    ct = Counters()
    secs = C.c_double()
    _ok(lib().orc_render_mt(C.byref(desc), C.c_uint64(seed), C.c_uint32(width), C.c_uint32(height),
                            C.c_uint64(first), C.c_uint64(n), C.c_uint64(batch), C.c_int(threads),
                            C.c_int(mode), _p(xyz), C.byref(ct), C.byref(secs)))
    return xyz, ct.as_dict(), float(secs.value)


def hardware_threads():
    return int(lib().orc_hardware_threads())
