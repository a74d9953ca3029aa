"""GPU parity: full-frame FAST (tracking.cpp:101) and cv::calcOpticalFlowPyrLK (tracking.cpp:593-618) through the C ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LK_TOL = 0.01  # px, BASELINE.json north_star: "LK tracked points must agree within 0.01 px, with an identical status mask"


@pytest.mark.parametrize("w,h,seed", [(1241, 376, 1000), (640, 200, 7), (333, 257, 11), (1920, 1080, 2000)])
@pytest.mark.parametrize("thr", [20, 7])
def test_fast_full_frame(pkg, orc, w, h, seed, thr):
    left, _ = pkg.synth_stereo_frame(seed, 2, w, h)
    ctx = pkg.Context(0, w, h, 2)
    g = ctx.fast_detect(left, thr)
    o = orc.fast9(left, thr)
    assert len(g) == len(o) and len(g) > 50
    for f in ("x", "y", "size", "angle", "response", "octave", "class_id"):
        assert np.array_equal(g[f], o[f]), f  # bit-exact, row-major order included
    ctx.close()


def test_fast_flat_and_tiny(pkg):
    ctx = pkg.Context(0, 640, 480, 2)
    assert len(ctx.fast_detect(np.full((480, 640), 9, np.uint8), 20)) == 0
    assert len(ctx.fast_detect(np.full((6, 6), 9, np.uint8), 20)) == 0  # smaller than the 7x7 FAST footprint
    ctx.close()


def _check_lk(g, o, n_min_ok):
    (gp, gs, ge), (op, os_, oe) = g, o
    assert np.array_equal(gs, os_), f"status differs at {np.nonzero(gs != os_)[0][:10]}"
    ok = os_ == 1
    assert ok.sum() >= n_min_ok
    d = np.abs(gp - op)[ok]
    assert d.max() <= LK_TOL, f"max |dp| = {d.max()}"
    assert np.abs(ge - oe)[ok].max() <= 1e-3
    return float(d.max())


@pytest.mark.parametrize("w,h,seed", [(1241, 376, 1000), (640, 200, 7), (333, 257, 11), (1920, 1080, 2000)])
def test_lk_stereo_and_temporal(pkg, orc, w, h, seed):
    l0, r0 = pkg.synth_stereo_frame(seed, 3, w, h)
    l1, r1 = pkg.synth_stereo_frame(seed, 4, w, h)
    ctx = pkg.Context(0, w, h, 2)
    kp = ctx.fast_detect(l0, 20)
    pts = np.stack([kp["x"], kp["y"]], 1)
    worst = 0.0
    for a, b in ((l0, r0), (r0, r1), (r1, l1), (l1, l0)):
        worst = max(worst, _check_lk(ctx.lk_track(a, b, pts), orc.lk_track(a, b, pts), len(pts) // 2))
    print(f"{w}x{h}: {len(pts)} points, max |dp| = {worst:.2e} px")
    ctx.close()


def test_lk_border_and_lost_points(pkg, orc):
    """Points on/over the image border, in flat regions (minEig), sub-pixel positions, and a scene change (lost tracks)."""
    w, h = 640, 200
    l0, r0 = pkg.synth_stereo_frame(21, 0, w, h)
    other, _ = pkg.synth_stereo_frame(99, 5, w, h)
    rng = np.random.default_rng(5)
    pts = np.concatenate([
        rng.uniform(-15, 30, (200, 2)), rng.uniform(0, 1, (300, 2)) * [w, h],
        np.stack([rng.uniform(w - 30, w + 15, 200), rng.uniform(h - 30, h + 15, 200)], 1),
        np.array([[0, 0], [w - 1, h - 1], [-11.5, -11.5], [w + 9.5, h + 9.5], [10.0, 10.0], [-31.0, 5.0], [5.0, -31.5]]),
    ]).astype(np.float32)
    ctx = pkg.Context(0, w, h, 2)
    flat = np.full((h, w), 100, np.uint8)
    flat[:, w // 2:] = l0[:, w // 2:]
    for a, b in ((l0, r0), (l0, other), (flat, flat), (l0, flat)):
        g = ctx.lk_track(a, b, pts)
        o = orc.lk_track(a, b, pts)
        assert np.array_equal(g[1], o[1])
        ok = o[1] == 1
        if ok.any():
            assert np.abs(g[0] - o[0])[ok].max() <= LK_TOL
    ctx.close()


def test_lk_levels_and_iterations(pkg, orc):
    w, h = 500, 120  # only 3 pyramid levels fit a 21-px window: 500x120, 250x60, 125x30 (63x15 is below win)
    l0, r0 = pkg.synth_stereo_frame(8, 1, w, h)
    ctx = pkg.Context(0, w, h, 2)
    kp = ctx.fast_detect(l0, 20)
    pts = np.stack([kp["x"], kp["y"]], 1)
    for max_level, max_iter, eps in ((3, 30, 0.01), (0, 30, 0.01), (1, 5, 0.03), (2, 1, 0.01)):
        g = ctx.lk_track(l0, r0, pts, max_level=max_level, max_iter=max_iter, eps=eps)
        o = orc.lk_track(l0, r0, pts, max_level=max_level, max_iter=max_iter, eps=eps)
        assert np.array_equal(g[1], o[1])
        ok = o[1] == 1
        assert np.abs(g[0] - o[0])[ok].max() <= LK_TOL
    ctx.close()


def test_lk_sharp_patterns_one_pixel_out_of_register(pkg, orc):
    """Binary stripes / checkerboards of period 4 shifted by one pixel: every window pixel carries the largest gradient (4080) and half
    of them the largest difference (8160), so the first Gauss-Newton step's sums sum(diff * Ix) reach ~7e9 -- beyond 32 bits.  The
    sums are exact integers in the oracle (64-bit); the device's must be too."""
    w, h = 320, 240
    yy, xx = np.mgrid[0:h, 0:w]
    stripes = (((xx >> 1) & 1) * 255).astype(np.uint8)
    checker = ((((xx >> 1) ^ (yy >> 1)) & 1) * 255).astype(np.uint8)
    rng = np.random.default_rng(3)
    pts = (rng.uniform(0, 1, (400, 2)) * [w, h]).astype(np.float32)
    ctx = pkg.Context(0, w, h, 2)
    big = 0
    for img in (stripes, checker):
        for sh in (1, -1):
            nxt = np.ascontiguousarray(np.roll(img, sh, axis=1))
            for max_level in (0, 3):
                g = ctx.lk_track(img, nxt, pts, max_level=max_level)
                o = orc.lk_track(img, nxt, pts, max_level=max_level)
                assert np.array_equal(g[1], o[1])
                ok = o[1] == 1
                big += int(ok.sum())
                if ok.any():
                    assert np.abs(g[0] - o[0])[ok].max() <= LK_TOL
                    assert np.abs(g[2] - o[2])[ok].max() <= 1e-3
    assert big > 400
    ctx.close()


@pytest.mark.parametrize("gap", [3, 6, 11])
def test_lk_large_motion_many_iterations(pkg, orc, gap):
    """Frames several steps apart (2 px of motion per step + the 8 px disparity): tracks need many Gauss-Newton steps, leave their
    integer cell repeatedly (tile re-staging), oscillate or get lost -- status identical, points within 0.01 px."""
    w, h = 800, 300
    l0, r0 = pkg.synth_stereo_frame(321, 1, w, h)
    l1, r1 = pkg.synth_stereo_frame(321, 1 + gap, w, h)
    ctx = pkg.Context(0, w, h, 2)
    kp = ctx.fast_detect(l0, 20)
    pts = np.stack([kp["x"], kp["y"]], 1) + np.float32(0.37)  # sub-pixel seeds
    for a, b in ((l0, l1), (l0, r1), (r1, l0)):
        g, o = ctx.lk_track(a, b, pts), orc.lk_track(a, b, pts)
        assert np.array_equal(g[1], o[1])
        ok = o[1] == 1
        assert ok.sum() > 20
        assert np.abs(g[0] - o[0])[ok].max() <= LK_TOL
        assert np.abs(g[2] - o[2])[ok].max() <= 1e-3
    ctx.close()


def test_gpu_against_cv2_golden_vectors(pkg):
    """The CUDA path directly against vectors made by the REAL OpenCV (cv2 4.13, oracle/make_golden.py): full-frame FAST and the
    the cv2.resize pyramid chain and the
    BF-Hamming match bit-exact, calcOpticalFlowPyrLK with identical status and within 0.01 px -- no oracle in between."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "cv2_primitives.npz"))
    l0, r0, l1 = g["left0"], g["right0"], g["left1"]
    h, w = l0.shape
    ctx = pkg.Context(0, w, h, 2)
    for thr in (20, 7):
        k = ctx.fast_detect(l0, thr)
        assert np.array_equal(np.stack([k["x"], k["y"], k["response"]], 1), g[f"fast_{thr}"])
    ctx.orb_configure(300, 1.2, 5, 20, 7)
    ctx.orb_extract(l0)
    for l in range(1, 5):  # cv2.resize chain exactly as ComputePyramid calls it
        assert np.array_equal(ctx.orb_read_level(0, l), g[f"resize_l{l}"]), l
    idx, dist = ctx.match_hamming(g["bf_q"], g["bf_t"])
    assert np.array_equal(idx, g["bf_idx"]) and np.array_equal(dist, g["bf_dist"])
    pts = g["fast_20"][:, :2].astype(np.float32)
    for nxt, key in ((r0, "lk_stereo"), (l1, "lk_temporal")):
        p, st, err = ctx.lk_track(l0, nxt, pts)
        assert np.array_equal(st, g[key + "_status"])
        ok = st == 1
        assert ok.sum() > 50 and np.abs(p - g[key + "_pts"])[ok].max() <= LK_TOL
        assert np.abs(err - g[key + "_err"])[ok].max() <= 1e-2
    ctx.close()
