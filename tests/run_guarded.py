"""Driven by test_robustness_gpu.py::test_no_kernel_writes_outside_its_buffers with SVO_GUARD=1: every entry point of the C ABI at
awkward sizes (odd widths, ring wrap-around, capacities reached), then the guard bands of all device buffers are inspected.
Prints one line: GUARD <overwritten bytes> <calls made>."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
assert pkg.guard_check() == 0, "SVO_GUARD=1 must be set before the library loads"
assert pkg.guard_selftest() == 3, "the guard bands must see a deliberate overrun (positive control)"
calls = 0
P1 = np.array([[718.856, 0, 607.1928, 0], [0, 718.856, 185.2157, 0], [0, 0, 1, 0]], np.float64)
P2 = P1.copy()
P2[0, 3] = -386.1448
for (w, h, nfeat, B, total) in ((333, 257, 300, 2, 5), (1241, 376, 1000, 3, 7), (1920, 1080, 2000, 2, 3), (97, 75, 50, 1, 2), (2051, 1023, 3000, 1, 2)):
    left, right = pkg.synth_sequence(77, 14, total, w, h)
    ctx = pkg.Context(0, w, h, 2 * (B + 1))
    ctx.orb_configure(nfeat, 1.2, 8, 20, 7)
    ctx.frontend_configure(w, h, pkg.MODE_ORB | pkg.MODE_LK, B)
    ctx.frontend_set_projection(P1, P2)
    ctx.frontend_set_pnp(True, 500 if w < 1500 else 70, 0.5)  # the pose step on both filtered lists of every frame (K13)
    kc, pc = ctx.orb_capacity(w, h), max(1024, w * h // 64)
    for f0 in range(0, total, B):  # several batches: the slot ring wraps
        n = min(B, total - f0)
        out = pkg.FrontendBuffers(n, kc, pc)
        ctx.frontend_batch(np.ascontiguousarray(left[f0:f0 + n]), np.ascontiguousarray(right[f0:f0 + n]), out)
        calls += 1
    # the standalone entry points on the same context
    ks, ds = ctx.orb_extract_batch([left[0], right[0]])
    k1, d1 = ctx.orb_extract(left[1])
    kp = ctx.fast_detect(left[0], 20)
    kp7 = ctx.fast_detect(left[0], 7)
    calls += 4
    if len(ds[0]) and len(ds[1]):
        i1, dd1 = ctx.match_hamming(ds[0], ds[1])
        i2, dd2 = ctx.match_hamming(ds[0], d1)
        ctx.match_hamming(ds[0][:1], ds[1])
        ctx.match_hamming(ds[0], ds[1][:1])
        ctx.orb_match_filter(ks[0], ks[1], k1, i1, dd1, i2, dd2, 3.0)
        calls += 5
    pts = np.stack([kp["x"], kp["y"]], 1)[:3000].astype(np.float32)
    edge = np.array([[0, 0], [w - 1, h - 1], [-10.5, 3], [w + 9.0, h + 9.0], [w - 0.5, 0.25]], np.float32)
    pts = np.concatenate([pts, edge])
    p1, s1, _ = ctx.lk_track(left[0], right[0], pts)
    p2, s2, _ = ctx.lk_track(right[0], right[1], p1)
    p3, s3, _ = ctx.lk_track(right[1], left[1], p2)
    p0r, s4, _ = ctx.lk_track(left[1], left[0], p3)
    ctx.lk_chain_filter(pts, p1, p2, p3, p0r, s1, s2, s3, s4, 3.0)
    _, xyz = ctx.triangulate_points(P1, P2, pts, p1)
    # the standalone pose call: one list, then a ragged batch whose longest list fills its capacity exactly
    ctx.pnp_ransac(P1[:, :3], xyz, p3, iterations=60)
    nb = min(len(xyz), 700)
    b3 = np.zeros((4, nb, 3), np.float32)
    b2 = np.zeros((4, nb, 2), np.float32)
    b3[:] = xyz[:nb]
    b2[:] = p3[:nb]
    ctx.pnp_ransac(P1[:, :3], b3, b2, [nb, 0, 5, nb // 2], iterations=40)
    calls += 8
    bad = pkg.guard_check()
    if bad:
        print(f"GUARD {bad} {calls} at {w}x{h}")
        sys.exit(1)
    ctx.close()
# brute-force matches at the sizes where the tile logic has its edges
ctx = pkg.Context(0, 64, 64, 2)
rng = np.random.default_rng(1)
for nq, nt in ((1, 1), (255, 63), (256, 64), (257, 65), (4000, 4000), (70000, 10), (10, 70000)):
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    ctx.match_hamming(q, t)
    calls += 1
bad = pkg.guard_check()
ctx.close()
print(f"GUARD {bad} {calls}")
sys.exit(0 if bad == 0 else 1)
