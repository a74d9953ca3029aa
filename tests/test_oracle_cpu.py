"""CPU suite: the oracle against the committed golden vectors (made by cv2, oracle/make_golden.py) and, when cv2 is
importable, against cv2 live on fresh inputs.  No GPU needed."""
import hashlib
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden", "cv2_primitives.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _kp3(k):
    return np.stack([k["x"], k["y"], k["response"]], 1)


def test_golden_resize_chain(orc, gold):
    cur = gold["left0"]
    for l in range(1, 5):
        ref = gold[f"resize_l{l}"]
        got = orc.resize_linear(cur, ref.shape[1], ref.shape[0])
        assert np.array_equal(got, ref), f"level {l}"
        cur = ref


def test_golden_fast(orc, gold):
    for thr in (20, 7):
        assert np.array_equal(_kp3(orc.fast9(gold["left0"], thr)), gold[f"fast_{thr}"])
    roi = np.ascontiguousarray(gold["left0"][16:16 + 37, 40:40 + 38])
    assert np.array_equal(_kp3(orc.fast9(roi, 7)), gold["fast_roi_7"])


def test_golden_blur_pyrdown_atan2_match(orc, gold):
    assert np.array_equal(orc.gaussian_blur7(gold["left0"]), gold["blur"])
    assert np.array_equal(orc.pyr_down(gold["left0"]), gold["pyrdown"])
    assert np.array_equal(orc.fast_atan2(gold["atan2_y"], gold["atan2_x"]), gold["atan2"])
    idx, dist = orc.match_hamming(gold["bf_q"], gold["bf_t"])
    assert np.array_equal(idx, gold["bf_idx"]) and np.array_equal(dist, gold["bf_dist"])
    assert idx[5] == 11  # duplicate train rows 11 and 40: the lowest index wins


def test_golden_lk(orc, gold):
    pts = gold["fast_20"][:, :2]
    for name, a, b in (("lk_stereo", "left0", "right0"), ("lk_temporal", "left0", "left1")):
        p, st, err = orc.lk_track(gold[a], gold[b], pts)
        assert np.array_equal(st, gold[name + "_status"])
        ok = st == 1
        assert np.abs(p - gold[name + "_pts"])[ok].max() <= 1e-3  # OpenCV sums in f32 SIMD lanes, the oracle in exact integers
        assert np.abs(err - gold[name + "_err"])[ok].max() <= 1e-2


def test_golden_triangulation(orc, gold):
    o4, o3 = orc.triangulate_points(gold["tri_P1"], gold["tri_P2"], gold["tri_pts1"], gold["tri_pts2"])
    ref3 = gold["tri_points3d"]
    assert (np.abs(o3 - ref3) <= 2e-5 * np.abs(ref3) + 1e-6).all()  # observed: bit-identical to cv2 4.13
    sgn = np.sign((o4 * gold["tri_points4d"]).sum(1))[:, None]  # the singular vector's sign is arbitrary
    assert np.abs(o4 * sgn - gold["tri_points4d"]).max() <= 1e-6


def test_tables(orc):
    scale, inv, fpl, umax = orc.orb_tables(1000, 1.2, 8)
    assert list(fpl) == [217, 181, 151, 126, 105, 87, 73, 60]
    assert list(orc.orb_tables(2000, 1.2, 8)[2]) == [434, 362, 302, 251, 209, 175, 145, 122]
    assert list(orc.orb_tables(4000, 1.2, 8)[2]) == [869, 724, 603, 503, 419, 349, 291, 242]
    assert list(umax) == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    assert scale[1] == np.float32(1.2) and inv[0] == 1.0


def test_brief_pattern_checksum(orc, pkg):
    pat = orc.brief_pattern()
    assert hashlib.sha256(pat.astype(np.int8).tobytes()).hexdigest() == "2164181aea6ff9ac426ca512d5130d15e1f6e3cd47b1cbdd568bbe1e55d49023"
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    a = open(os.path.join(here, "oracle", "brief_pattern.inc")).read()
    b = open(os.path.join(here, "stereo-visual-odometry_b200", "csrc", "brief_pattern.inc")).read()
    assert a == b  # oracle and product carry the same table


def test_orb_extract_structure(orc, pkg):
    """ORBextractor restatement: level-major order, quotas, borders, descriptor determinism."""
    left, _ = pkg.synth_stereo_frame(1000, 0, 1241, 376)
    o = orc.OrbOracle(1000)
    k, d = o.extract(left)
    assert 1000 <= len(k) <= 1000 + 3 * 8 and d.shape == (len(k), 32)
    assert (np.diff(k["octave"]) >= 0).all()
    _, _, fpl, _ = orc.orb_tables(1000, 1.2, 8)
    for l in range(8):
        lk = o.level_keypoints(l)
        assert fpl[l] <= len(lk) <= fpl[l] + 3
        lvl = o.level(l)
        assert (lk["x"] >= 19).all() and (lk["x"] < lvl.shape[1] - 19).all() and (lk["y"] >= 19).all() and (lk["y"] < lvl.shape[0] - 19).all()
        c = o.candidates(l)
        cs = {(x, y) for x, y in zip(c["x"], c["y"])}
        assert all((x - 16, y - 16) in cs for x, y in zip(lk["x"], lk["y"]))  # every selected keypoint is a candidate
    k2, d2 = orc.OrbOracle(1000).extract(left)
    assert np.array_equal(k, k2) and np.array_equal(d, d2)
    kf, df = o.extract(np.full((376, 1241), 50, np.uint8))
    assert len(kf) == 0


def test_octree_properties(orc):
    rng = np.random.default_rng(3)
    for trial in range(30):
        W, H = int(rng.integers(60, 1300)), int(rng.integers(40, 400))
        if W < H // 2 + 1:
            continue
        n = int(rng.integers(0, 3000))
        pos = rng.permutation(W * H)[:n]
        keys = np.zeros(n, orc.KP_DTYPE)
        keys["x"], keys["y"] = pos % W, pos // W
        keys["response"] = rng.integers(7, 120, n)
        N = int(rng.integers(1, 400))
        out, st = orc.distribute_octree(keys, 16, 16 + W, 16, 16 + H, N)
        nini = int(np.round(np.float32(W) / np.float32(H)))
        assert len(out) <= max(N + 3, 4 * max(nini, 1))
        assert len(out) == len({(a, b) for a, b in zip(out["x"], out["y"])})  # one key per node, nodes are disjoint
        if n >= N * 6:
            assert len(out) >= N
        if n:
            assert len(out) >= 1


def test_match_and_lk_filters(orc):
    rng = np.random.default_rng(0)
    n = 50
    ll = np.zeros(n, orc.KP_DTYPE)
    lr = np.zeros(n, orc.KP_DTYPE)
    ll["y"] = rng.integers(0, 100, n)
    lr["y"] = ll["y"] + rng.integers(-5, 6, n)
    idx1 = np.arange(n, dtype=np.int32)
    d1 = rng.integers(0, 80, n).astype(np.int32)
    d2 = rng.integers(0, 80, n).astype(np.int32)
    acc = orc.orb_match_filter(idx1, d1, idx1, d2, n, n - 5, ll, lr, 3.0)  # des_index = n - 5 truncates the loop (quirk Q3)
    mind = np.maximum(d1, d2)[:n - 5].min()
    lim = max(2 * mind, 30)
    exp = [i for i in range(n - 5) if d1[i] <= lim and d2[i] <= lim and abs(ll["y"][i] - lr["y"][i]) < 3]
    assert list(acc[:, 0]) == exp
    p = rng.uniform(-2, 50, (5, n, 2)).astype(np.float32)
    s = (rng.uniform(0, 1, (4, n)) > 0.1).astype(np.uint8)
    keep = orc.lk_chain_filter(*p, *s, 3.0)
    exp = [(p[:, i] >= 0).all() and s[:, i].all() and abs(p[0, i, 1] - p[1, i, 1]) <= 3 and abs(p[2, i, 1] - p[3, i, 1]) <= 3 for i in range(n)]
    assert list(keep) == exp


def test_live_cv2_cross_check(orc, pkg):
    cv2 = pytest.importorskip("cv2")
    cv2.setNumThreads(1)
    left, right = pkg.synth_stereo_frame(77, 3, 500, 300)
    assert np.array_equal(orc.resize_linear(left, 417, 250), cv2.resize(left, (417, 250), interpolation=cv2.INTER_LINEAR))
    assert np.array_equal(orc.gaussian_blur7(left), cv2.GaussianBlur(left, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101))
    kp = cv2.FastFeatureDetector_create(20, True).detect(left)
    assert np.array_equal(_kp3(orc.fast9(left, 20)), np.array([(k.pt[0], k.pt[1], k.response) for k in kp], np.float32).reshape(-1, 3))
    pts = np.array([k.pt for k in kp], np.float32)
    p, st, _ = cv2.calcOpticalFlowPyrLK(left, right, pts.reshape(-1, 1, 2), None, winSize=(21, 21), maxLevel=3,
                                        criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01), flags=0, minEigThreshold=1e-3)
    op, ost, _ = orc.lk_track(left, right, pts)
    assert np.array_equal(ost, st.ravel()) and np.abs(op - p.reshape(-1, 2))[ost == 1].max() <= 1e-3
    # the per-cell FAST loop of the extractor against cv2.FAST called per cell (ORBextractor.cpp:742-781)
    o = orc.OrbOracle(500)
    o.extract(left)
    lvl = o.level(1)
    H, W = lvl.shape
    maxBX, maxBY = W - 16, H - 16
    nC, nR = int(np.float32(maxBX - 16) / np.float32(30)), int(np.float32(maxBY - 16) / np.float32(30))
    wC, hC = int(np.ceil(np.float32(maxBX - 16) / nC)), int(np.ceil(np.float32(maxBY - 16) / nR))
    f20, f7 = cv2.FastFeatureDetector_create(20, True), cv2.FastFeatureDetector_create(7, True)
    exp = []
    for i in range(nR):
        y0, y1 = 16 + i * hC, min(16 + i * hC + hC + 6, maxBY)
        if y0 >= maxBY - 3:
            continue
        for j in range(nC):
            x0, x1 = 16 + j * wC, min(16 + j * wC + wC + 6, maxBX)
            if x0 >= maxBX - 6:
                continue
            kk = f20.detect(lvl[y0:y1, x0:x1]) or f7.detect(lvl[y0:y1, x0:x1])
            exp += [(k.pt[0] + j * wC, k.pt[1] + i * hC, k.response) for k in kk]
    assert np.array_equal(_kp3(o.candidates(1)), np.array(exp, np.float32).reshape(-1, 3))


def test_cv2_primitive_orb_equals_oracle(pkg, orc):
    """CPU baseline variant 2 (real cv2.resize / cv2.FAST per cell / cv2.GaussianBlur + the oracle's octree / IC_Angle / BRIEF loops)
    gives the oracle's keypoints and descriptors: the three baseline variants time the same computation."""
    cv2 = pytest.importorskip("cv2")
    import ref_pipeline
    for (w, h, nfeat, seed) in ((640, 200, 500, 7), (1241, 376, 1000, 1000)):
        left, _ = pkg.synth_stereo_frame(seed, 2, w, h)
        k0, d0 = orc.OrbOracle(nfeat, 1.2, 8, 20, 7).extract(left)
        k1, d1 = ref_pipeline.Cv2PrimitiveOrb(nfeat, 1.2, 8, 20, 7).extract(left)
        assert len(k0) == len(k1) and np.array_equal(k0, k1)
        assert np.array_equal(d0, d1)


# ---- pose from points (oracle/oracle_pnp.cpp) against cv2 4.13 outputs (oracle/make_golden_pnp.py) ----------------------------
@pytest.fixture(scope="module")
def gold_pnp():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "cv2_pnp.npz"))


def test_pnp_svd_matches_cv2(orc, gold_pnp):
    g = gold_pnp
    for i in range(6):
        for kind in ("svd3", "svdg", "svd12"):
            w, u, vt = orc.svd(g[f"{kind}_{i}_a"])
            scale = g[f"{kind}_{i}_w"].max()
            assert np.abs(w - g[f"{kind}_{i}_w"]).max() <= 1e-12 * scale, (kind, i)
            assert np.abs(u - g[f"{kind}_{i}_u"]).max() <= 1e-9, (kind, i)   # signs and order included
            if kind != "svd12":
                assert np.abs(vt - g[f"{kind}_{i}_vt"]).max() <= 1e-9, (kind, i)


def _rod(orc, r):
    R = np.zeros(9)
    orc.lib().orc_rodrigues(orc._p(np.ascontiguousarray(r, np.float64)), orc._p(R))
    return R.reshape(3, 3)


def test_pnp_epnp_matches_cv2(orc, gold_pnp):
    g = gold_pnp
    for k in range(int(g["epnp_count"])):
        ok, R, t = orc.epnp(g["K"], g[f"epnp_{k}_p3"], g[f"epnp_{k}_p2"])
        assert ok
        assert np.abs(R - _rod(orc, g[f"epnp_{k}_r"])).max() <= 1e-9, k
        assert np.abs(t - g[f"epnp_{k}_t"]).max() <= 1e-9, k
    # the minimal case: the two-dimensional null space of M^T M makes about one case in seven depend on rounding noise inside
    # the SVD (oracle_pnp.cpp header); the rest must agree
    same = 0
    for k in range(40):
        ok, R, t = orc.epnp(g["K"], g[f"epnp5_{k}_p3"], g[f"epnp5_{k}_p2"])
        same += bool(ok and np.abs(R - _rod(orc, g[f"epnp5_{k}_r"])).max() <= 1e-6 and np.abs(t - g[f"epnp5_{k}_t"]).max() <= 1e-6)
    assert same >= 30, same


def test_pnp_refine_matches_cv2(orc, gold_pnp):
    g = gold_pnp
    for k in range(6):
        r, t, it = orc.pnp_refine(g["K"], g[f"lm_{k}_p3"], g[f"lm_{k}_p2"], np.zeros(3), np.zeros(3))
        assert np.abs(r - g[f"lm_{k}_r"]).max() <= 1e-8 and np.abs(t - g[f"lm_{k}_t"]).max() <= 1e-8, k
        assert 1 <= it <= 20


def test_pnp_ransac_matches_cv2(orc, gold_pnp):
    g = gold_pnp
    n_scenes = int(g["ransac_count"])
    exact = 0
    for k in range(n_scenes):
        res = orc.pnp_ransac(g["K"], g[f"ransac_{k}_p3"], g[f"ransac_{k}_p2"])
        m = g[f"ransac_{k}_mask"].astype(bool)
        assert res["ok"] == bool(g[f"ransac_{k}_ok"])
        jac = (m & res["mask"]).sum() / max((m | res["mask"]).sum(), 1)
        dp = max(np.abs(res["rvec"] - g[f"ransac_{k}_r"]).max(), np.abs(res["tvec"] - g[f"ransac_{k}_t"]).max())
        if np.array_equal(m, res["mask"]):
            exact += 1
            assert dp <= 1e-8, (k, dp)       # same consensus set -> same refined pose
        else:
            assert jac >= 0.8 and dp <= 1e-2, (k, jac, dp)
    assert exact >= n_scenes // 2, (exact, n_scenes)   # identical inlier set AND pose on the majority of the scenes


def test_pnp_degenerate_inputs(orc, gold_pnp):
    K = gold_pnp["K"]
    res = orc.pnp_ransac(K, np.zeros((3, 3), np.float32), np.zeros((3, 2), np.float32))
    assert not res["ok"] and res["n_inliers"] == 0
    p3, p2 = gold_pnp["ransac_0_p3"][:5], gold_pnp["ransac_0_p2"][:5]
    orc.pnp_ransac(K, p3, p2)  # n == model size: one hypothesis from points 0..4, no crash
    res = orc.pnp_ransac(K, np.zeros((50, 3), np.float32), np.zeros((50, 2), np.float32))  # all points identical: nothing usable
    assert not res["ok"] or res["n_inliers"] >= 5


def test_pnp_ransac_live_cv2(orc, gold_pnp):
    """Fresh scenes against cv2 itself (when importable): the same statement as the fixture test on other seeds and sizes."""
    cv2 = pytest.importorskip("cv2")
    cv2.setNumThreads(1)
    K = gold_pnp["K"]
    rng = np.random.default_rng(99)
    exact, n_scenes = 0, 24
    for k in range(n_scenes):
        n = int(rng.integers(30, 900))
        P = np.stack([rng.uniform(-15, 15, n), rng.uniform(-3, 3, n), rng.uniform(5, 50, n)], 1).astype(np.float32)
        rv, tv = rng.normal(0, 0.02, 3), np.array([rng.normal(0, .1), rng.normal(0, .05), -rng.uniform(0.3, 1.5)])
        Xc = P.astype(np.float64) @ cv2.Rodrigues(rv)[0].T + tv
        uv = (Xc[:, :2] / Xc[:, 2:]) * [K[0, 0], K[1, 1]] + [K[0, 2], K[1, 2]] + rng.normal(0, 0.15, (n, 2))
        no = int(rng.uniform(0.05, 0.5) * n)
        uv[:no] += rng.uniform(-30, 30, (no, 2))
        uv = uv.astype(np.float32)
        ok, r, t, inl = cv2.solvePnPRansac(P, uv, K, np.zeros((4, 1)), rvec=np.zeros((3, 1)), tvec=np.zeros((3, 1)), useExtrinsicGuess=True,
                                           iterationsCount=500, reprojectionError=0.5, confidence=0.99, flags=cv2.SOLVEPNP_ITERATIVE)
        m = np.zeros(n, bool)
        if inl is not None:
            m[inl.ravel()] = True
        res = orc.pnp_ransac(K, P, uv)
        assert res["ok"] == bool(ok)
        if np.array_equal(m, res["mask"]):
            exact += 1
            assert max(np.abs(res["rvec"] - r.ravel()).max(), np.abs(res["tvec"] - t.ravel()).max()) <= 1e-8, k
        else:
            assert (m & res["mask"]).sum() / max((m | res["mask"]).sum(), 1) >= 0.8, k
    assert exact >= (3 * n_scenes) // 4, (exact, n_scenes)
