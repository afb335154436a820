"""The measurement scripts under tools/ and bench.py need a GPU to run; at least they must parse."""
import glob
import os
import py_compile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_scripts_compile(tmp_path):
    scripts = sorted(glob.glob(os.path.join(ROOT, "tools", "*.py"))) + [os.path.join(ROOT, "bench.py"),
                                                                        os.path.join(ROOT, "__graft_entry__.py")]
    assert len(scripts) >= 8
    for k, path in enumerate(scripts):
        py_compile.compile(path, cfile=str(tmp_path / f"{k}.pyc"), doraise=True)
