"""The lazy host buffers of the C++ unit mirror (host/rl_units.hpp), checked on the CPU."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_mirror_semantics(pkg, tmp_path):
    exe = str(tmp_path / "host_mirror_check")
    pkg_dir = os.path.dirname(pkg.LIB_PATH)
    subprocess.run(["g++", "-O1", "-std=c++17", "-Wall", "-Wextra", os.path.join(ROOT, "tests", "cpp", "host_mirror_check.cpp"),
                    "-o", exe, "-L" + pkg_dir, "-lrl_b200", "-Wl,-rpath," + pkg_dir], check=True)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert res.returncode == 0, res.stderr
    assert "host mirror ok" in res.stdout
