"""The C-ABI library loads and exports exactly what include/rl_b200.h declares;
without a GPU every compute entry point fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header="rl_b200.h"):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rl_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(pkg):
    handle = C.CDLL(pkg.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 45
    missing = [n for n in names if not hasattr(handle, n)]
    assert not missing, f"declared in rl_b200.h but not exported: {missing}"
    # the host-only scene builders live in a library of their own, without CUDA in it
    host = C.CDLL(pkg.HOST_LIB_PATH)
    host_names = declared_symbols("rl_host.h")
    assert len(host_names) >= 13
    missing = [n for n in host_names if not hasattr(host, n)]
    assert not missing, f"declared in rl_host.h but not exported: {missing}"
    assert not [n for n in host_names if hasattr(handle, n)], "scene builders leaked into the product library"
    import subprocess
    needed = subprocess.run(["readelf", "-d", pkg.HOST_LIB_PATH], capture_output=True, text=True).stdout
    assert "cuda" not in needed.lower()


def test_python_mirror_binds_every_declared_symbol(pkg):
    assert sorted(pkg.SYMBOLS) == declared_symbols()
    assert sorted(pkg.HOST_SYMBOLS) == declared_symbols("rl_host.h")
    pkg.lib()
    pkg.host_lib()
    assert pkg.lib().rl_abi_version() == 3


def test_pod_layouts_match_reference_records(pkg):
    # MappedPhoton = 4 x f32 (trace_unit.rs:23-37); Vector3 = 3 x f32 (vector3.rs:20-25)
    assert pkg.MAPPED_PHOTON.itemsize == 16
    assert C.sizeof(pkg.Vec3) == 12 and C.sizeof(pkg.Quat) == 16
    assert C.sizeof(pkg.Surface) == 4 + 36 + 4 + 8
    assert C.sizeof(pkg.Material) == 16 and C.sizeof(pkg.Object) == 20
    assert pkg.RAY.itemsize == 32 and pkg.HIT.itemsize == 44
    assert pkg.BATCH_PHOTONS == 524288


def test_builtin_scene_descriptor(pkg):
    # app.rs:166-363: 1 sun + 3 paraboloids + 2 sky circles + ceiling + 100 + 100 + 110 spheres + 22 prisms
    d = pkg.SceneBuilder(pkg.SCENE_C2).desc()
    assert d.n_objects == 339
    kinds = [d.surfaces[d.objects[i].surface].kind for i in range(d.n_objects)]
    assert kinds.count(pkg.SURFACE_SPHERE) == 311
    assert kinds.count(pkg.SURFACE_PARABOLOID) == 3
    assert kinds.count(pkg.SURFACE_CIRCLE) == 2
    assert kinds.count(pkg.SURFACE_PLANE) == 1
    assert kinds.count(pkg.SURFACE_COMPOUND) == 22
    mats = [d.objects[i].material.kind for i in range(d.n_objects)]
    assert mats.count(pkg.MATERIAL_BLACKBODY) == 3
    assert mats.count(pkg.MATERIAL_SOAP_BUBBLE) == 110
    assert mats.count(pkg.MATERIAL_GLOSSY_MIRROR) == 100
    assert mats.count(pkg.MATERIAL_SF10_GLASS) == 22
    # hexagonal prism = 8 half-spaces + 7 compound nodes (geometry.rs:409-416)
    n_half = sum(1 for i in range(d.n_surfaces) if d.surfaces[i].kind == pkg.SURFACE_HALFSPACE)
    assert n_half == 22 * 8
    # first seed sphere: i = 19 (app.rs:237), radius 0.8
    first = d.surfaces[d.objects[7].surface]
    assert first.kind == pkg.SURFACE_SPHERE and abs(first.s - 0.64) < 1e-6
    assert d.camera.kind == pkg.CAMERA_ORBIT


def test_other_scene_descriptors(pkg):
    assert pkg.SceneBuilder(pkg.SCENE_C1).desc().n_objects == 2
    assert pkg.SceneBuilder(pkg.SCENE_C3).desc().n_objects == 3
    assert pkg.SceneBuilder(pkg.SCENE_C4).desc().n_objects == 4098
    assert pkg.SceneBuilder(pkg.SCENE_C4, 64).desc().n_objects == 66


def test_no_cpu_fallback(pkg):
    if pkg.device_count() > 0:
        pytest.skip("a GPU is present")
    b = pkg.SceneBuilder(pkg.SCENE_C1)
    with pytest.raises(pkg.RlError) as e:
        pkg.Scene(b)
    assert e.value.code == pkg.RL_ERR_CUDA
    for ctor in (lambda: pkg.TraceUnit(0, 8, 8), lambda: pkg.PlotUnit(0, 8, 8),
                 lambda: pkg.GatherUnit(8, 8), lambda: pkg.TonemapUnit(8, 8)):
        with pytest.raises(pkg.RlError) as e:
            ctor()
        assert e.value.code == pkg.RL_ERR_CUDA
    with pytest.raises(pkg.RlError):
        pkg.debug_math(0, np.zeros(4, np.float32))


def test_product_does_not_reference_the_oracle():
    # the oracle is test infrastructure: nothing under the package may import,
    # include or dlopen it
    pkg_dir = os.path.join(ROOT, "robigo-luculenta_b200")
    for base, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                text = open(os.path.join(base, f), errors="replace").read()
                assert "oracle" not in text.lower(), f"{f} mentions the oracle"


def test_header_is_plain_c(tmp_path):
    # the boundary is a C ABI: the header compiles as C99, pedantically, without a C++ compiler
    import subprocess
    src = tmp_path / "use_header.c"
    src.write_text('#include "rl_host.h"\n'
                   "int use(void) {\n"
                   "    rl_scene_desc d; rl_mapped_photon p; rl_hit h; rl_ray r;\n"
                   "    (void)d; (void)p; (void)h; (void)r;\n"
                   "    return (int)sizeof(rl_surface) + RL_ABI_VERSION + (int)RL_BATCH_PHOTONS + (rl_abi_version() == RL_ABI_VERSION);\n"
                   "}\n")
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    "-c", str(src), "-o", str(tmp_path / "use_header.o")], check=True)
