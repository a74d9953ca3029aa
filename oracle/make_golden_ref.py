#!/usr/bin/env python
"""Generates tests/golden/ref_orbextractor.npz: outputs of the REFERENCE'S OWN ORBextractor.cpp, compiled from
/root/reference/src/ORBextractor.cpp against oracle/ref_shim (make -C oracle ref), run under the monotone allocator
(oracle/ref_capi.cpp) on a few synthetic images.  The vectors pin the oracle restatement (CPU tests) and the CUDA path (GPU
tests) where /root/reference and oracle/_ref do not exist.  Run from the repo root in the build container:
    python oracle/make_golden_ref.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

# (key, seed, frame, w, h, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)
CASES = [
    ("kitti", 1000, 0, 1241, 376, 1000, 1.2, 8, 20, 7),
    ("small", 7, 0, 640, 200, 500, 1.2, 8, 20, 7),
    ("odd", 11, 0, 333, 257, 300, 1.2, 8, 20, 7),
    ("thr", 4242, 3, 752, 480, 800, 1.2, 6, 30, 10),
]


def main():
    pkg, orc = g.load_package(), g.load_oracle()
    assert orc.ref_lib() is not None, "build oracle/_ref first (needs /root/reference)"
    out = {"cases": np.array([repr(c) for c in CASES])}
    for key, seed, frame, w, h, nf, sf, nl, ini, mn in CASES:
        img, _ = pkg.synth_stereo_frame(seed, frame, w, h)
        k, d = orc.RefOrb(nf, sf, nl, ini, mn, monotone=True).extract(img)
        out[key + "_kps"], out[key + "_desc"] = k, d
        print(key, len(k), "keypoints")
    path = os.path.join(ROOT, "tests", "golden", "ref_orbextractor.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
