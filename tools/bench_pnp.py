"""K13 alone on scenes with depth (the bench sequence is a fronto-parallel plane, the worst case for EPnP): 128 lists per launch,
ORB-sized (n = 1000) and LK-sized (n = 5000), 30 % / 60 % gross outliers, 0.15-px noise.  Wall time of svo_pnp_ransac() includes
its host<->device copies; kernel time comes from an ncu launch list of this script (tools/gpu_round.sh)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

K = np.array([[718.856, 0, 607.1928], [0, 718.856, 185.2157], [0, 0, 1]])


def rodrigues(r):
    th = np.linalg.norm(r)
    k = r / th
    kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.cos(th) * np.eye(3) + (1 - np.cos(th)) * np.outer(k, k) + np.sin(th) * kx


def scenes(B, n, outl, rng):
    p3 = np.zeros((B, n, 3), np.float32)
    p2 = np.zeros((B, n, 2), np.float32)
    for b in range(B):
        P = np.stack([rng.uniform(-15, 15, n), rng.uniform(-3, 3, n), rng.uniform(5, 50, n)], 1).astype(np.float32)
        rv = rng.normal(0, 0.02, 3)
        tv = np.array([rng.normal(0, .1), rng.normal(0, .05), -rng.uniform(0.3, 1.5)])
        Xc = P.astype(np.float64) @ rodrigues(rv).T + tv
        uv = (Xc[:, :2] / Xc[:, 2:]) * [K[0, 0], K[1, 1]] + [K[0, 2], K[1, 2]] + rng.normal(0, 0.15, (n, 2))
        uv[:int(outl * n)] += rng.uniform(-30, 30, (int(outl * n), 2))
        p3[b], p2[b] = P, uv
    return p3, p2


def main():
    pkg = g.load_package()
    ctx = pkg.Context(0, 64, 64, 2)
    rng = np.random.default_rng(5)
    for n, outl in ((1000, 0.3), (5000, 0.3), (1000, 0.6), (5000, 0.6)):
        p3, p2 = scenes(128, n, outl, rng)
        ctx.pnp_ransac(K, p3, p2)
        t = time.perf_counter()
        for _ in range(3):
            r = ctx.pnp_ransac(K, p3, p2)
        dt = (time.perf_counter() - t) / 3
        print(f"n {n} outliers {outl}: {dt * 1e3:.2f} ms per 128 lists (incl. copies), hypotheses evaluated median {int(np.median(r['iters']))} "
              f"max {int(r['iters'].max())}, inliers median {int(np.median(r['n_inliers']))}, ok {int(r['ok'].sum())}/128")
    ctx.close()


if __name__ == "__main__":
    main()
