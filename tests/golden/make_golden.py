#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the oracle.

These are REGRESSION PINS OF THE ORACLE in specified-math mode, not outputs of
the reference: the reference is Rust with an unseeded RNG and cannot be built
or run in this environment (see oracle/oracle.cpp header: parity unpinned).
They freeze today's restatement so that later edits to the oracle or to the
CUDA path cannot drift silently; the pins that tie the oracle to the reference
source are the known-answer tests in tests/test_oracle_kat.py.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import __graft_entry__ as entry  # noqa: E402
import oracle_lib as orc  # noqa: E402

CASES = {          # name: (builtin scene, param, width, height, seed, first photon, photons)
    "c1_sphere_plane": (1, 0, 256, 256, 0x5EED, 0, 512),
    "c2_builtin": (2, 0, 1024, 1024, 0x5EED, 0, 512),
    "c2_builtin_high_ids": (2, 0, 1280, 720, 7, (1 << 36) + 5, 256),
    "c3_prism": (3, 0, 1024, 1024, 0x5EED, 0, 512),
    "c4_spheres_64": (4, 64, 512, 512, 0x5EED, 0, 256),
}


def main():
    entry.build_library()
    entry.build_oracle()
    pkg = entry.load_package()
    for name, (which, param, w, h, seed, first, n) in CASES.items():
        desc = pkg.SceneBuilder(which, param).desc()
        ct = orc.Counters()
        photons = orc.trace(desc, seed, w, h, first, n, orc.MATH_SPEC, False, ct)
        rays, _ = orc.camera_rays(desc, seed, w, h, first, 32)
        hits = orc.intersect(desc, rays)
        image = orc.plot(32, 32 * h // w if h <= w else 32, photons)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), photons=photons, rays=np.uint64(ct.rays),
                            camera_rays=rays, hits=hits, image=image,
                            meta=np.array([which, param, w, h, seed, first, n], dtype=np.uint64))
        print(name, "rays", ct.rays, "lit", int(np.count_nonzero(photons["probability"])))


if __name__ == "__main__":
    main()
