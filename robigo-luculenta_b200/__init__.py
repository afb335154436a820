"""robigo-luculenta_b200 -- Python mirror of the reference's unit API over the C ABI.

The classes keep the reference's names, methods and public fields
(`TraceUnit.render` / `.mapped_photons`, `PlotUnit.plot` / `.clear` /
`.tristimulus_buffer`, `GatherUnit.accumulate` / `.save` / `.tristimulus_buffer`,
`TonemapUnit.tonemap` / `.rgb_buffer`; reference: src/trace_unit.rs:56-168,
src/plot_unit.rs:34-102, src/gather_unit.rs:26-92, src/tonemap_unit.rs:30-100)
so that tests read like calls into the Rust units.  Everything computes on the
GPU through `librl_b200.so` (include/rl_b200.h); there is no CPU fallback and
importing this module fails loudly if the library has not been built.

The directory name carries a hyphen, so the package is imported under the name
`robigo_luculenta_b200` via `__graft_entry__.load_package()`.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# RL_B200_LIB overrides the library path (tuning experiments build variants beside the default)
LIB_PATH = os.environ.get("RL_B200_LIB") or os.path.join(_HERE, "librl_b200.so")
# the host-only scene builders (include/rl_host.h): no CUDA in it
HOST_LIB_PATH = os.path.join(_HERE, "librl_host.so")

BATCH_PHOTONS = 1024 * 512        # trace_unit.rs:67
TEST_BATCH_PHOTONS = 1024         # trace_unit.rs:70

# ----------------------------------------------------------------- ctypes PODs


class Vec3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]


class Quat(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("w", C.c_float)]


class Surface(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("a", Vec3), ("b", Vec3), ("c", Vec3), ("s", C.c_float),
                ("child", C.c_uint32 * 2)]


class Material(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("p0", C.c_float), ("p1", C.c_float), ("p2", C.c_float)]


class Object(C.Structure):
    _fields_ = [("surface", C.c_uint32), ("material", Material)]


class Camera(C.Structure):
    _fields_ = [("position", Vec3), ("field_of_view", C.c_float), ("focal_distance", C.c_float),
                ("depth_of_field", C.c_float), ("chromatic_abberation", C.c_float),
                ("orientation", Quat)]


class CameraModel(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("fixed", Camera),
                ("phi_base", C.c_float), ("phi_rate", C.c_float),
                ("alpha_base", C.c_float), ("alpha_rate", C.c_float),
                ("distance_base", C.c_float), ("distance_rate", C.c_float),
                ("focal_factor", C.c_float),
                ("keyframes", C.POINTER(Camera)), ("n_keyframes", C.c_uint32)]


class SceneDesc(C.Structure):
    _fields_ = [("surfaces", C.POINTER(Surface)), ("n_surfaces", C.c_uint32),
                ("objects", C.POINTER(Object)), ("n_objects", C.c_uint32),
                ("camera", CameraModel)]


MAPPED_PHOTON = np.dtype([("x", "<f4"), ("y", "<f4"), ("probability", "<f4"), ("wavelength", "<f4")])
RAY = np.dtype([("origin", "<f4", 3), ("direction", "<f4", 3), ("wavelength", "<f4"),
                ("probability", "<f4")])
HIT = np.dtype([("object", "<i4"), ("distance", "<f4"), ("position", "<f4", 3),
                ("normal", "<f4", 3), ("tangent", "<f4", 3)])

SURFACE_PLANE, SURFACE_HALFSPACE, SURFACE_CIRCLE, SURFACE_SPHERE, SURFACE_PARABOLOID, SURFACE_COMPOUND = range(1, 7)
(MATERIAL_BLACKBODY, MATERIAL_DIFFUSE_GREY, MATERIAL_DIFFUSE_COLOURED, MATERIAL_GLOSSY_MIRROR,
 MATERIAL_SF10_GLASS, MATERIAL_SOAP_BUBBLE) = range(1, 7)
CAMERA_STATIC, CAMERA_ORBIT, CAMERA_KEYFRAMES = 1, 2, 3
SCENE_C1, SCENE_C2, SCENE_C3, SCENE_C4 = 1, 2, 3, 4
SCENE_C6 = 6        # compounds over spheres and half-spaces

RL_OK, RL_ERR_INVALID, RL_ERR_UNSUPPORTED, RL_ERR_CUDA, RL_ERR_IO, RL_ERR_NOMEM = 0, -1, -2, -3, -4, -5


class RlError(RuntimeError):
    """A non-zero status from the C ABI (the Rust shim `expect`s these)."""

    def __init__(self, code, message):
        super().__init__(f"rl_b200 status {code}: {message}")
        self.code = code


# Every symbol include/rl_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_U64, _U32, _I = C.c_uint64, C.c_uint32, C.c_int
_PF = C.POINTER(C.c_float)
SYMBOLS = {
    "rl_abi_version": (_I, []),
    "rl_last_error": (C.c_char_p, []),
    "rl_device_count": (_I, []),
    "rl_kernel_launch_count": (_U64, []),
    "rl_kernel_launch_count_reset": (None, []),
    "rl_scene_create": (_I, [C.POINTER(SceneDesc), C.POINTER(_P)]),
    "rl_scene_destroy": (_I, [_P]),
    "rl_trace_unit_create": (_I, [_U64, _U32, _U32, _U64, C.POINTER(_P)]),
    "rl_trace_unit_destroy": (_I, [_P]),
    "rl_trace_unit_set_batch_size": (_I, [_P, _U64]),
    "rl_trace_unit_set_stream": (_I, [_P, _P]),
    "rl_trace_unit_render": (_I, [_P, _P, _P]),
    "rl_trace_unit_render_async": (_I, [_P, _P, _P]),
    "rl_trace_unit_download": (_I, [_P, _P, _U64, C.POINTER(_U64)]),
    "rl_trace_unit_render_range": (_I, [_P, _P, _U64, _U64, _P]),
    "rl_trace_unit_render_fused": (_I, [_P, _P, _P, _U64, _U64]),
    "rl_trace_unit_ray_count": (_I, [_P, C.POINTER(_U64)]),
    "rl_trace_unit_sync": (_I, [_P]),
    "rl_scene_batch_counter_reset": (_I, [_P, _U64]),
    "rl_scene_dispatch_stats": (_I, [_P, C.POINTER(_U64), C.POINTER(_U64)]),
    "rl_transfer_counters": (None, [C.POINTER(_U64), C.POINTER(_U64)]),
    "rl_transfer_counters_reset": (None, []),
    "rl_host_register": (_I, [_P, C.c_size_t]),
    "rl_host_unregister": (_I, [_P]),
    "rl_plot_unit_create": (_I, [_U64, _U32, _U32, C.POINTER(_P)]),
    "rl_plot_unit_destroy": (_I, [_P]),
    "rl_plot_unit_set_stream": (_I, [_P, _P]),
    "rl_plot_unit_plot": (_I, [_P, _P, _U64]),
    "rl_plot_unit_plot_device": (_I, [_P, _P]),
    "rl_plot_unit_clear": (_I, [_P]),
    "rl_plot_unit_download": (_I, [_P, _P]),
    "rl_plot_unit_device_buffer": (_I, [_P, C.POINTER(_P), C.POINTER(C.c_size_t)]),
    "rl_plot_unit_sync": (_I, [_P]),
    "rl_plot_unit_ipc_export": (_I, [_P, _P]),
    "rl_ipc_open": (_I, [_P, C.POINTER(_P)]),
    "rl_ipc_close": (_I, [_P]),
    "rl_gather_unit_create": (_I, [_U32, _U32, C.c_char_p, C.POINTER(_P)]),
    "rl_gather_unit_destroy": (_I, [_P]),
    "rl_gather_unit_set_stream": (_I, [_P, _P]),
    "rl_gather_unit_accumulate": (_I, [_P, _P]),
    "rl_gather_unit_accumulate_plot": (_I, [_P, _P, _I]),
    "rl_gather_unit_accumulate_device": (_I, [_P, C.POINTER(_P), _U32]),
    "rl_gather_unit_save": (_I, [_P, C.c_char_p]),
    "rl_gather_unit_set_save_interval": (_I, [_P, C.c_double]),
    "rl_gather_unit_flush": (_I, [_P]),
    "rl_gather_unit_load": (_I, [_P, C.c_char_p]),
    "rl_gather_unit_download": (_I, [_P, _P, _P]),
    "rl_gather_unit_sync": (_I, [_P]),
    "rl_tonemap_unit_create": (_I, [_U32, _U32, C.POINTER(_P)]),
    "rl_tonemap_unit_destroy": (_I, [_P]),
    "rl_tonemap_unit_set_stream": (_I, [_P, _P]),
    "rl_tonemap_unit_tonemap": (_I, [_P, _P, _P]),
    "rl_tonemap_unit_tonemap_gather": (_I, [_P, _P, _P]),
    "rl_tonemap_unit_set_exposure_mode": (_I, [_P, _I]),
    "rl_tonemap_unit_last_exposure": (_I, [_P, _PF]),
    "rl_debug_intersect": (_I, [_P, _P, _U64, _P]),
    "rl_debug_math": (_I, [_I, _P, _P, _U64, _P]),
    "rl_debug_tristimulus": (_I, [_P, _U64, _P]),
    "rl_debug_camera_rays": (_I, [_P, _U64, _U32, _U32, _U64, _U64, _P, _P]),
    "rl_debug_cull_check": (_I, [_P, _U64, _U32, _U32, _U64, _U64, C.POINTER(_U64), C.POINTER(_U64)]),
}

# Every symbol include/rl_host.h declares (librl_host.so)
HOST_SYMBOLS = {
    "rl_scene_builder_create": (_I, [C.POINTER(_P)]),
    "rl_scene_builder_destroy": (_I, [_P]),
    "rl_scene_builder_builtin": (_I, [_P, _I, _U32]),
    "rl_scene_builder_plane": (_I, [_P, Vec3, Vec3]),
    "rl_scene_builder_circle": (_I, [_P, Vec3, Vec3, C.c_float]),
    "rl_scene_builder_sphere": (_I, [_P, Vec3, C.c_float]),
    "rl_scene_builder_halfspace": (_I, [_P, Vec3, Vec3]),
    "rl_scene_builder_compound": (_I, [_P, _U32, _U32]),
    "rl_scene_builder_paraboloid": (_I, [_P, Vec3, Vec3, C.c_float]),
    "rl_scene_builder_prism": (_I, [_P, Vec3, Vec3, C.c_float, C.c_float, C.c_float]),
    "rl_scene_builder_hexagonal_prism": (_I, [_P, Vec3, Vec3, C.c_float, C.c_float, C.c_float, C.c_float]),
    "rl_material_blackbody": (_I, [C.c_float, C.c_float, C.POINTER(Material)]),
    "rl_scene_builder_object": (_I, [_P, _U32, Material]),
    "rl_scene_builder_camera": (_I, [_P, C.POINTER(CameraModel)]),
    "rl_scene_builder_desc": (_I, [_P, C.POINTER(SceneDesc)]),
}

_lib_handle = None
_host_handle = None


def lib():
    """The C-ABI library; raises if it has not been built (no fallback)."""
    global _lib_handle
    if _lib_handle is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
                "robigo-luculenta_b200 has no CPU fallback")
        handle = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SYMBOLS.items():
            fn = getattr(handle, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib_handle = handle
    return _lib_handle


def host_lib():
    """The host-only scene builders; raises if the library has not been built."""
    global _host_handle
    if _host_handle is None:
        if not os.path.exists(HOST_LIB_PATH):
            raise ImportError(f"{HOST_LIB_PATH} is missing: build it with "
                              "`python -c 'import __graft_entry__ as g; g.build()'`")
        handle = C.CDLL(HOST_LIB_PATH)
        for name, (restype, argtypes) in HOST_SYMBOLS.items():
            fn = getattr(handle, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _host_handle = handle
    return _host_handle


def _check(code):
    if code != RL_OK:
        raise RlError(code, lib().rl_last_error().decode("utf-8", "replace"))


def _check_host(code):
    if code != RL_OK:
        raise RlError(code, "scene builder (librl_host.so)")


def _ptr(array):
    return None if array is None else array.ctypes.data_as(C.c_void_p)


def device_count():
    return lib().rl_device_count()


def kernel_launch_count():
    return int(lib().rl_kernel_launch_count())


def reset_kernel_launch_count():
    lib().rl_kernel_launch_count_reset()


def transfer_counters():
    """(host->device, device->host) bytes copied through the C ABI since the last reset."""
    a, b = _U64(0), _U64(0)
    lib().rl_transfer_counters(C.byref(a), C.byref(b))
    return int(a.value), int(b.value)


def reset_transfer_counters():
    lib().rl_transfer_counters_reset()




def vec3(x, y, z):
    return Vec3(float(x), float(y), float(z))


# ------------------------------------------------------------- scene building
class SceneBuilder:
    """Host-side scene description (mirror of App::set_up_scene and the
    geometry/material constructors; app.rs:166-363, geometry.rs, material.rs)."""

    def __init__(self, builtin=None, param=0):
        self._h = _P()
        _check_host(host_lib().rl_scene_builder_create(C.byref(self._h)))
        if builtin is not None:
            _check_host(host_lib().rl_scene_builder_builtin(self._h, builtin, param))

    def __del__(self):
        if getattr(self, "_h", None) and lib is not None:
            host_lib().rl_scene_builder_destroy(self._h)
            self._h = None

    def _idx(self, r):
        if r < 0:
            raise RlError(r, "scene builder")
        return r

    def plane(self, normal, offset):
        return self._idx(host_lib().rl_scene_builder_plane(self._h, vec3(*normal), vec3(*offset)))

    def circle(self, normal, position, radius):
        return self._idx(host_lib().rl_scene_builder_circle(self._h, vec3(*normal), vec3(*position), radius))

    def sphere(self, position, radius):
        return self._idx(host_lib().rl_scene_builder_sphere(self._h, vec3(*position), radius))

    def halfspace(self, normal, offset):
        return self._idx(host_lib().rl_scene_builder_halfspace(self._h, vec3(*normal), vec3(*offset)))

    def compound(self, surface1, surface2):
        return self._idx(host_lib().rl_scene_builder_compound(self._h, surface1, surface2))

    def paraboloid(self, normal, offset, focal_distance):
        return self._idx(host_lib().rl_scene_builder_paraboloid(self._h, vec3(*normal), vec3(*offset), focal_distance))

    def prism(self, axis, offset, edge_length, angle, height):
        return self._idx(host_lib().rl_scene_builder_prism(self._h, vec3(*axis), vec3(*offset), edge_length, angle, height))

    def hexagonal_prism(self, axis, offset, edge_length, bevel_size, angle, height):
        return self._idx(host_lib().rl_scene_builder_hexagonal_prism(
            self._h, vec3(*axis), vec3(*offset), edge_length, bevel_size, angle, height))

    @staticmethod
    def blackbody(kelvins, intensity):
        m = Material()
        _check_host(host_lib().rl_material_blackbody(kelvins, intensity, C.byref(m)))
        return m

    @staticmethod
    def material(kind, p0=0.0, p1=0.0, p2=0.0):
        return Material(kind, p0, p1, p2)

    def object(self, surface, material):
        return self._idx(host_lib().rl_scene_builder_object(self._h, surface, material))

    def static_camera(self, position, orientation=(0.0, 0.0, 0.0, 1.0), field_of_view=0.35 * np.pi,
                      focal_distance=1.0, depth_of_field=1.0e9, chromatic_abberation=0.0):
        cm = CameraModel()
        cm.kind = CAMERA_STATIC
        cm.fixed = Camera(vec3(*position), field_of_view, focal_distance, depth_of_field,
                          chromatic_abberation, Quat(*[float(v) for v in orientation]))
        _check_host(host_lib().rl_scene_builder_camera(self._h, C.byref(cm)))

    def camera_model(self, cm):
        _check_host(host_lib().rl_scene_builder_camera(self._h, C.byref(cm)))

    def keyframe_camera(self, cameras):
        """A tabulated `fn(f32) -> Camera` (scene.rs:34): camera(t) = cameras[min(floor(t n), n - 1)].
        `cameras`: Camera structures, or (position, orientation, fov, focal, dof, ca) tuples."""
        frames = (Camera * len(cameras))()
        for k, c in enumerate(cameras):
            frames[k] = c if isinstance(c, Camera) else Camera(vec3(*c[0]), c[2], c[3], c[4], c[5],
                                                             Quat(*[float(v) for v in c[1]]))
        cm = CameraModel()
        cm.kind = CAMERA_KEYFRAMES
        cm.fixed = frames[0]
        cm.keyframes = frames
        cm.n_keyframes = len(cameras)
        _check_host(host_lib().rl_scene_builder_camera(self._h, C.byref(cm)))

    def desc(self):
        """Borrowed descriptor (valid while the builder is alive and unchanged)."""
        d = SceneDesc()
        _check_host(host_lib().rl_scene_builder_desc(self._h, C.byref(d)))
        d._owner = self
        return d


class Scene:
    """Device-resident scene (stands in for Arc<Scene>, app.rs:63)."""

    def __init__(self, desc_or_builder):
        self.builder = desc_or_builder if isinstance(desc_or_builder, SceneBuilder) else None
        self.desc = desc_or_builder.desc() if self.builder else desc_or_builder
        self._h = _P()
        _check(lib().rl_scene_create(C.byref(self.desc), C.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None) and lib is not None:
            lib().rl_scene_destroy(self._h)
            self._h = None

    def reset_batch_counter(self, next_batch=0):
        """The next TraceUnit.render on this scene takes batch number `next_batch`."""
        _check(lib().rl_scene_batch_counter_reset(self._h, next_batch))

    def dispatch_stats(self):
        """(launches, batches) of the scene's trace dispatcher: how many TraceUnit.render
        batches went out with how many launches."""
        a, b = _U64(0), _U64(0)
        _check(lib().rl_scene_dispatch_stats(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    # probes ---------------------------------------------------------------
    def intersect(self, rays):
        rays = np.ascontiguousarray(rays, dtype=RAY)
        out = np.zeros(rays.shape[0], dtype=HIT)
        _check(lib().rl_debug_intersect(self._h, _ptr(rays), rays.shape[0], _ptr(out)))
        return out

    def cull_check(self, seed, width, height, first, n):
        """(rays compared, rays where the culled intersect differed from brute force)."""
        rays, bad = _U64(0), _U64(0)
        _check(lib().rl_debug_cull_check(self._h, seed, width, height, first, n, C.byref(rays), C.byref(bad)))
        return int(rays.value), int(bad.value)

    def camera_rays(self, seed, width, height, first, n):
        rays = np.zeros(n, dtype=RAY)
        xy = np.zeros(n, dtype=MAPPED_PHOTON)
        _check(lib().rl_debug_camera_rays(self._h, seed, width, height, first, n, _ptr(rays), _ptr(xy)))
        return rays, xy


# ---------------------------------------------------------------------- units
class TraceUnit:
    """trace_unit.rs:40-168"""

    def __init__(self, id, width, height, seed=0x5EED, batch=BATCH_PHOTONS):
        self.id, self.width, self.height, self.seed = id, width, height, seed
        self._h = _P()
        _check(lib().rl_trace_unit_create(id, width, height, seed, C.byref(self._h)))
        self.batch = batch
        _check(lib().rl_trace_unit_set_batch_size(self._h, batch))
        self.mapped_photons = np.zeros(0, dtype=MAPPED_PHOTON)

    def __del__(self):
        if getattr(self, "_h", None) and lib is not None:
            lib().rl_trace_unit_destroy(self._h)
            self._h = None

    def set_stream(self, cuda_stream):
        _check(lib().rl_trace_unit_set_stream(self._h, _P(cuda_stream)))

    def render(self, scene, download=True, wait=True):
        """TraceUnit::render (trace_unit.rs:151-168); fills `mapped_photons`.  With
        wait=False the batch is only queued (rl_trace_unit_render_async): call sync()
        before reading `mapped_photons`."""
        if download:
            if self.mapped_photons.shape[0] != self.batch:
                self.mapped_photons = np.zeros(self.batch, dtype=MAPPED_PHOTON)
            fn = lib().rl_trace_unit_render if wait else lib().rl_trace_unit_render_async
            _check(fn(self._h, scene._h, _ptr(self.mapped_photons)))
        else:
            _check(lib().rl_trace_unit_render(self._h, scene._h, None))

    def download(self, out=None):
        """Records of the last render, copied out of the device now (rl_trace_unit_download)."""
        if out is None:
            out = np.zeros(self.batch, dtype=MAPPED_PHOTON)
        count = _U64(0)
        _check(lib().rl_trace_unit_download(self._h, _ptr(out), out.shape[0], C.byref(count)))
        out = out[: int(count.value)]
        self.mapped_photons = out
        return out

    def render_range(self, scene, first_photon, n_photons, download=True, out=None):
        if download:
            if out is None:
                out = np.zeros(n_photons, dtype=MAPPED_PHOTON)
            _check(lib().rl_trace_unit_render_range(self._h, scene._h, first_photon, n_photons, _ptr(out)))
            self.mapped_photons = out
            return out
        _check(lib().rl_trace_unit_render_range(self._h, scene._h, first_photon, n_photons, None))
        return None

    def render_fused(self, scene, plot_unit, first_photon, n_photons):
        _check(lib().rl_trace_unit_render_fused(self._h, scene._h, plot_unit._h, first_photon, n_photons))

    def ray_count(self):
        v = _U64(0)
        _check(lib().rl_trace_unit_ray_count(self._h, C.byref(v)))
        return int(v.value)

    def sync(self):
        _check(lib().rl_trace_unit_sync(self._h))


class PlotUnit:
    """plot_unit.rs:23-103"""

    def __init__(self, id, width, height):
        self.id, self.width, self.height = id, width, height
        self._h = _P()
        _check(lib().rl_plot_unit_create(id, width, height, C.byref(self._h)))
        self._buffer = None

    def __del__(self):
        if getattr(self, "_h", None) and lib is not None:
            lib().rl_plot_unit_destroy(self._h)
            self._h = None

    def set_stream(self, cuda_stream):
        _check(lib().rl_plot_unit_set_stream(self._h, _P(cuda_stream)))

    def plot(self, photons):
        """PlotUnit::plot (plot_unit.rs:87-95): a host slice, or a TraceUnit
        whose records are still on the device."""
        if isinstance(photons, TraceUnit):
            _check(lib().rl_plot_unit_plot_device(self._h, photons._h))
        else:
            photons = np.ascontiguousarray(photons, dtype=MAPPED_PHOTON)
            _check(lib().rl_plot_unit_plot(self._h, _ptr(photons), photons.shape[0]))

    def clear(self):
        _check(lib().rl_plot_unit_clear(self._h))

    def download(self, out=None):
        if out is None:
            out = np.zeros((self.height, self.width, 3), dtype=np.float32)
        _check(lib().rl_plot_unit_download(self._h, _ptr(out)))
        return out

    @property
    def tristimulus_buffer(self):
        return self.download()

    def device_buffer(self):
        p, n = _P(), C.c_size_t()
        _check(lib().rl_plot_unit_device_buffer(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def sync(self):
        _check(lib().rl_plot_unit_sync(self._h))

    def ipc_export(self):
        """64-byte handle another process can open with ipc_open()."""
        buf = C.create_string_buffer(64)
        _check(lib().rl_plot_unit_ipc_export(self._h, buf))
        return buf.raw


class GatherUnit:
    """gather_unit.rs:24-94"""

    def __init__(self, width, height, resume_path=None):
        self.width, self.height = width, height
        self._h = _P()
        path = None if resume_path is None else os.fsencode(resume_path)
        _check(lib().rl_gather_unit_create(width, height, path, C.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None) and lib is not None:
            lib().rl_gather_unit_destroy(self._h)
            self._h = None

    def set_stream(self, cuda_stream):
        _check(lib().rl_gather_unit_set_stream(self._h, _P(cuda_stream)))

    def accumulate(self, tristimuli, clear=False):
        """GatherUnit::accumulate (gather_unit.rs:49-64): a host buffer or a PlotUnit."""
        if isinstance(tristimuli, PlotUnit):
            _check(lib().rl_gather_unit_accumulate_plot(self._h, tristimuli._h, 1 if clear else 0))
        else:
            t = np.ascontiguousarray(tristimuli, dtype=np.float32)
            assert t.size == self.width * self.height * 3
            _check(lib().rl_gather_unit_accumulate(self._h, _ptr(t)))

    def accumulate_device(self, pointers):
        arr = (_P * len(pointers))(*[_P(p) for p in pointers])
        _check(lib().rl_gather_unit_accumulate_device(self._h, arr, len(pointers)))

    def save(self, path="buffer.raw", wait=True):
        """GatherUnit::save (gather_unit.rs:68-78).  The file is written by the unit's
        writer thread; `wait` returns once it is on disk (rl_gather_unit_flush)."""
        _check(lib().rl_gather_unit_save(self._h, os.fsencode(path)))
        if wait:
            self.flush()

    def set_save_interval(self, seconds):
        """At most one background buffer.raw write is started per interval (default 0.1 s)."""
        _check(lib().rl_gather_unit_set_save_interval(self._h, float(seconds)))

    def flush(self):
        _check(lib().rl_gather_unit_flush(self._h))

    def load(self, path="buffer.raw"):
        _check(lib().rl_gather_unit_load(self._h, os.fsencode(path)))

    def download(self, with_compensation=False, out=None):
        if out is None:
            out = np.zeros((self.height, self.width, 3), dtype=np.float32)
        comp = np.zeros_like(out) if with_compensation else None
        _check(lib().rl_gather_unit_download(self._h, _ptr(out), _ptr(comp)))
        return (out, comp) if with_compensation else out

    @property
    def tristimulus_buffer(self):
        return self.download()

    def sync(self):
        _check(lib().rl_gather_unit_sync(self._h))


class TonemapUnit:
    """tonemap_unit.rs:22-101"""

    def __init__(self, width, height):
        self.width, self.height = width, height
        self._h = _P()
        _check(lib().rl_tonemap_unit_create(width, height, C.byref(self._h)))
        self.rgb_buffer = np.zeros((height, width, 3), dtype=np.uint8)

    def __del__(self):
        if getattr(self, "_h", None) and lib is not None:
            lib().rl_tonemap_unit_destroy(self._h)
            self._h = None

    def set_stream(self, cuda_stream):
        _check(lib().rl_tonemap_unit_set_stream(self._h, _P(cuda_stream)))

    def tonemap(self, tristimuli, download=True):
        """TonemapUnit::tonemap (tonemap_unit.rs:73-100): a host buffer or a GatherUnit
        (`download=False`: leave the image on the device)."""
        if isinstance(tristimuli, GatherUnit):
            _check(lib().rl_tonemap_unit_tonemap_gather(self._h, tristimuli._h,
                                                         _ptr(self.rgb_buffer) if download else None))
        else:
            t = np.ascontiguousarray(tristimuli, dtype=np.float32)
            assert t.size == self.width * self.height * 3
            _check(lib().rl_tonemap_unit_tonemap(self._h, _ptr(t), _ptr(self.rgb_buffer)))
        return self.rgb_buffer

    def set_exposure_mode(self, reference_fold):
        """find_exposure by the reference's sequential f32 folds (True) or the f64 reduction (False)."""
        _check(lib().rl_tonemap_unit_set_exposure_mode(self._h, 1 if reference_fold else 0))

    @property
    def last_exposure(self):
        v = C.c_float()
        _check(lib().rl_tonemap_unit_last_exposure(self._h, C.byref(v)))
        return float(v.value)


def ipc_open(handle_bytes):
    """Device pointer of a peer process's accumulator (see PlotUnit.ipc_export)."""
    p = _P()
    _check(lib().rl_ipc_open(handle_bytes, C.byref(p)))
    return p.value


def ipc_close(ptr):
    _check(lib().rl_ipc_close(_P(ptr)))


# --------------------------------------------------------------- math probes
def debug_math(fn, x, x2=None):
    x = np.ascontiguousarray(x, dtype=np.float32)
    x2 = None if x2 is None else np.ascontiguousarray(x2, dtype=np.float32)
    out = np.zeros_like(x)
    _check(lib().rl_debug_math(fn, _ptr(x), _ptr(x2), x.size, _ptr(out)))
    return out


def debug_tristimulus(wavelengths):
    w = np.ascontiguousarray(wavelengths, dtype=np.float32)
    out = np.zeros((w.size, 3), dtype=np.float32)
    _check(lib().rl_debug_tristimulus(_ptr(w), w.size, _ptr(out)))
    return out
