#!/usr/bin/env python
"""Generates tests/golden/cv2_pnp.npz: outputs of the REAL OpenCV (cv2 4.13.0) for the pose step -- cv2.SVDecomp on the matrix
sizes EPnP uses, cv2.solvePnP(SOLVEPNP_EPNP), cv2.solvePnP(SOLVEPNP_ITERATIVE, useExtrinsicGuess=True) and
cv2.solvePnPRansac called exactly as Tracking::OpenCV_EstimatePose_PnP calls it (tracking.cpp:486-488: rvec = tvec = 0 as the
guess, 500 iterations, 0.5 px, 0.99, SOLVEPNP_ITERATIVE) -- on seeded synthetic scenes (KITTI-like intrinsics, points 5-50 m
ahead, a small frame-to-frame motion, pixel noise and gross outliers).  Run from the repo root:
    python oracle/make_golden_pnp.py
"""
import os

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
K = np.array([[718.856, 0, 607.1928], [0, 718.856, 185.2157], [0, 0, 1]])


def scene(n, noise, outl, rng):
    P = np.stack([rng.uniform(-15, 15, n), rng.uniform(-3, 3, n), rng.uniform(5, 50, n)], 1).astype(np.float32)
    rv = rng.normal(0, 0.02, 3)
    tv = np.array([rng.normal(0, .1), rng.normal(0, .05), -rng.uniform(0.3, 1.5)])
    R = cv2.Rodrigues(rv)[0]
    Xc = P.astype(np.float64) @ R.T + tv
    uv = (Xc[:, :2] / Xc[:, 2:]) * [K[0, 0], K[1, 1]] + [K[0, 2], K[1, 2]]
    uv += rng.normal(0, noise, uv.shape)
    no = int(outl * n)
    uv[:no] += rng.uniform(-30, 30, (no, 2))
    return P, uv.astype(np.float32)


def main():
    cv2.setNumThreads(1)
    rng = np.random.default_rng(20261018)
    out = {"cv2_version": np.array(cv2.__version__), "K": K}
    # SVD: symmetric positive (semi-)definite 3 x 3 and full-rank 12 x 12 inputs, plus general 3 x 3
    for i in range(6):
        P = rng.normal(0, 1, (rng.integers(5, 30), 3)) * rng.uniform(0.1, 10, 3)
        P0 = P - P.mean(0)
        a = P0.T @ P0
        w, u, vt = cv2.SVDecomp(a)
        out[f"svd3_{i}_a"], out[f"svd3_{i}_w"], out[f"svd3_{i}_u"], out[f"svd3_{i}_vt"] = a, w.ravel(), u, vt
        g = rng.normal(0, 3, (3, 3))
        w, u, vt = cv2.SVDecomp(g)
        out[f"svdg_{i}_a"], out[f"svdg_{i}_w"], out[f"svdg_{i}_u"], out[f"svdg_{i}_vt"] = g, w.ravel(), u, vt
        M = rng.normal(0, 1, (14, 12)) * rng.uniform(0.1, 100, 12)
        a = M.T @ M
        w, u, vt = cv2.SVDecomp(a)
        out[f"svd12_{i}_a"], out[f"svd12_{i}_w"], out[f"svd12_{i}_u"] = a, w.ravel(), u
    # EPnP, n >= 6
    k = 0
    for n in (6, 7, 10, 50):
        for noise in (0.0, 0.3):
            for _ in range(5):
                P, uv = scene(n, noise, 0, rng)
                ok, r, t = cv2.solvePnP(P, uv, K, None, flags=cv2.SOLVEPNP_EPNP)
                out[f"epnp_{k}_p3"], out[f"epnp_{k}_p2"], out[f"epnp_{k}_r"], out[f"epnp_{k}_t"] = P, uv, r.ravel(), t.ravel()
                k += 1
    out["epnp_count"] = np.array(k)
    # EPnP, n = 5 (the minimal case; agreement is counted, not required per case)
    for k in range(40):
        P, uv = scene(5, 0.3, 0, rng)
        ok, r, t = cv2.solvePnP(P, uv, K, None, flags=cv2.SOLVEPNP_EPNP)
        out[f"epnp5_{k}_p3"], out[f"epnp5_{k}_p2"], out[f"epnp5_{k}_r"], out[f"epnp5_{k}_t"] = P, uv, r.ravel(), t.ravel()
    # LM refinement from the zero guess
    for k in range(6):
        P, uv = scene(200, 0.3, 0, rng)
        ok, r, t = cv2.solvePnP(P, uv, K, None, rvec=np.zeros((3, 1)), tvec=np.zeros((3, 1)), useExtrinsicGuess=True, flags=cv2.SOLVEPNP_ITERATIVE)
        out[f"lm_{k}_p3"], out[f"lm_{k}_p2"], out[f"lm_{k}_r"], out[f"lm_{k}_t"] = P, uv, r.ravel(), t.ravel()
    # the whole call
    k = 0
    for n, outl in ((100, 0.1), (300, 0.2), (600, 0.3), (1000, 0.4), (400, 0.6), (40, 0.2)):
        for _ in range(6):
            P, uv = scene(n, 0.15, outl, rng)
            ok, r, t, inl = cv2.solvePnPRansac(P, uv, K, np.zeros((4, 1)), rvec=np.zeros((3, 1)), tvec=np.zeros((3, 1)), useExtrinsicGuess=True,
                                               iterationsCount=500, reprojectionError=0.5, confidence=0.99, flags=cv2.SOLVEPNP_ITERATIVE)
            m = np.zeros(n, np.uint8)
            if inl is not None:
                m[inl.ravel()] = 1
            out[f"ransac_{k}_p3"], out[f"ransac_{k}_p2"] = P, uv
            out[f"ransac_{k}_r"], out[f"ransac_{k}_t"], out[f"ransac_{k}_mask"], out[f"ransac_{k}_ok"] = r.ravel(), t.ravel(), m, np.array(ok)
            k += 1
    out["ransac_count"] = np.array(k)
    path = os.path.join(ROOT, "tests", "golden", "cv2_pnp.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
