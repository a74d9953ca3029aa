#!/usr/bin/env python
"""Generates tests/golden/cv2_primitives.npz: outputs of the REAL OpenCV primitives (cv2) on a small fixed input.

This is how the oracle is pinned: the reference ships no golden vectors, so the vectors are made
from the third-party dependency it calls (OpenCV; this image has cv2 4.13.0).  Run from the repo root:
    python oracle/make_golden.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402


def main():
    pkg = g.load_package()
    cv2.setNumThreads(1)
    w, h = 320, 200
    l0, r0 = pkg.synth_stereo_frame(424242, 0, w, h)
    l1, _ = pkg.synth_stereo_frame(424242, 1, w, h)
    out = {"cv2_version": np.array(cv2.__version__), "left0": l0, "right0": r0, "left1": l1}
    # resize chain exactly as ComputePyramid calls it (ORBextractor.cpp:1066,1074)
    s, cur = np.float32(1.0), l0
    for l in range(1, 5):
        s = np.float32(s * np.float32(1.2))
        inv = np.float32(1.0) / s
        dw, dh = int(np.rint(np.float32(w) * inv)), int(np.rint(np.float32(h) * inv))
        cur = cv2.resize(cur, (dw, dh), interpolation=cv2.INTER_LINEAR)
        out[f"resize_l{l}"] = cur
    for thr in (20, 7):
        kp = cv2.FastFeatureDetector_create(thr, True).detect(l0)
        out[f"fast_{thr}"] = np.array([(k.pt[0], k.pt[1], k.response) for k in kp], np.float32).reshape(-1, 3)
    roi = l0[16:16 + 37, 40:40 + 38]  # a cell-sized ROI view (non-contiguous), as in ORBextractor.cpp:763
    kp = cv2.FastFeatureDetector_create(7, True).detect(roi)
    out["fast_roi_7"] = np.array([(k.pt[0], k.pt[1], k.response) for k in kp], np.float32).reshape(-1, 3)
    out["blur"] = cv2.GaussianBlur(l0, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    out["pyrdown"] = cv2.pyrDown(l0)
    rng = np.random.default_rng(7)
    ay = rng.integers(-50000, 50000, 2000).astype(np.float32)
    ax = rng.integers(-50000, 50000, 2000).astype(np.float32)
    ay[:4], ax[:4] = [0, 0, 1, -1], [0, 5, 0, 0]
    out["atan2_y"], out["atan2_x"] = ay, ax
    out["atan2"] = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in zip(ay, ax)], np.float32)
    q = rng.integers(0, 256, (150, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (170, 32), dtype=np.uint8)
    t[40] = t[11]
    q[5] = t[11]
    m = cv2.BFMatcher(cv2.NORM_HAMMING).match(q, t)
    out["bf_q"], out["bf_t"] = q, t
    out["bf_idx"] = np.array([x.trainIdx for x in m], np.int32)
    out["bf_dist"] = np.array([x.distance for x in m], np.int32)
    pts = out["fast_20"][:, :2].copy()
    lk = dict(winSize=(21, 21), maxLevel=3, criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01), flags=0, minEigThreshold=1e-3)
    for name, (a, b) in {"lk_stereo": (l0, r0), "lk_temporal": (l0, l1)}.items():
        p, st, err = cv2.calcOpticalFlowPyrLK(a, b, pts.reshape(-1, 1, 2), None, **lk)
        out[name + "_pts"], out[name + "_status"], out[name + "_err"] = p.reshape(-1, 2), st.ravel(), err.ravel()
    # cv::triangulatePoints + convertPointsFromHomogeneous with KITTI-00-like projection matrices (sensors.cpp:8-24)
    P1 = np.array([[718.856, 0, 607.1928, 0], [0, 718.856, 185.2157, 0], [0, 0, 1, 0]], np.float64)
    P2 = P1.copy()
    P2[0, 3] = -386.1448
    xl = rng.uniform(20, 1220, 400).astype(np.float32)
    yl = rng.uniform(20, 356, 400).astype(np.float32)
    disp = rng.uniform(1.5, 90, 400).astype(np.float32)
    t1 = np.stack([xl, yl], 1)
    t2 = np.stack([xl - disp, yl + rng.uniform(-1.5, 1.5, 400).astype(np.float32)], 1)
    p4 = cv2.triangulatePoints(P1, P2, t1.T.copy(), t2.T.copy())
    p3 = cv2.convertPointsFromHomogeneous(p4.T.copy()).reshape(-1, 3)
    out["tri_P1"], out["tri_P2"], out["tri_pts1"], out["tri_pts2"] = P1, P2, t1, t2
    out["tri_points4d"], out["tri_points3d"] = p4.T.astype(np.float32).copy(), p3.astype(np.float32)
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    path = os.path.join(ROOT, "tests", "golden", "cv2_primitives.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()
