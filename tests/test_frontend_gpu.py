"""GPU parity of the throughput mode: batches of consecutive stereo frames with the previous frame carried on the device."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _oracle_chain(orc, l0, r0, l1, r1, seeds):
    p1, s1, _ = orc.lk_track(l0, r0, seeds)
    p2, s2, _ = orc.lk_track(r0, r1, p1)
    p3, s3, _ = orc.lk_track(r1, l1, p2)
    p0r, s4, _ = orc.lk_track(l1, l0, p3)
    return (p1, p2, p3, p0r), (s1, s2, s3, s4)


def _check_unit(pkg, orc, oracle, out, i, lf, rf, prev):
    """One unit of a batch (stereo frame i of `out`, images lf/rf) against the oracle; `prev` describes the frame before it (None for the
    first of a sequence).  Returns the `prev` of the next frame."""
    okl, odl = oracle.extract(lf)
    okr, odr = oracle.extract(rf)
    nl, nr = out.n_left[i], out.n_right[i]
    assert nl == len(okl) and nr == len(okr)
    assert np.array_equal(out.kps_left[i, :nl], okl) and np.array_equal(out.kps_right[i, :nr], okr)
    assert np.unpackbits(out.desc_left[i, :nl] ^ odl).sum() <= 2
    assert np.unpackbits(out.desc_right[i, :nr] ^ odr).sum() <= 2
    ofast = orc.fast9(lf, 20)
    assert out.n_fast[i] == len(ofast) and np.array_equal(out.fast_kps[i, :len(ofast)], ofast)
    if prev is None:
        assert out.n_query[i] == 0 and out.n_seed[i] == 0
    else:
        # K6 on the descriptors the GPU produced for the previous / current frame
        nq = out.n_query[i]
        assert nq == len(prev["dl"])
        si, sd = orc.match_hamming(prev["dl"], prev["dr"])
        ti, td = orc.match_hamming(prev["dl"], out.desc_left[i, :nl])
        assert np.array_equal(out.stereo_idx[i, :nq], si) and np.array_equal(out.stereo_dist[i, :nq], sd)
        assert np.array_equal(out.temporal_idx[i, :nq], ti) and np.array_equal(out.temporal_dist[i, :nq], td)
        # the 4-call LK chain seeded with the previous frame's FAST corners
        ns = out.n_seed[i]
        seeds = np.stack([prev["fast"]["x"], prev["fast"]["y"]], 1)
        assert ns == len(seeds)
        (p1, p2, p3, p0r), sts = _oracle_chain(orc, prev["l"], prev["r"], lf, rf, seeds)
        cum = np.ones(ns, bool)
        for k, (gp, op) in enumerate(zip((out.lk_p1, out.lk_p2, out.lk_p3, out.lk_p0r), (p1, p2, p3, p0r))):
            gs = out.lk_status[i, k, :ns]
            # cumulative status: a failed point carries garbage into the next call (quirk Q5)
            assert np.array_equal(gs[cum], sts[k][cum]), f"status of call {k + 1}"
            cum &= sts[k] == 1
            assert np.abs(gp[i, :ns] - op)[cum].max() <= 0.01, f"points of call {k + 1}"
        keep_g = orc.lk_chain_filter(seeds, out.lk_p1[i, :ns], out.lk_p2[i, :ns], out.lk_p3[i, :ns], out.lk_p0r[i, :ns],
                                     *[out.lk_status[i, k, :ns] for k in range(4)])
        keep_o = orc.lk_chain_filter(seeds, p1, p2, p3, p0r, *sts)
        assert np.array_equal(keep_g, keep_o)
        # K11: deleteBadmatchFeatures on the device == the oracle filter applied to the device's own chain outputs
        nk = out.n_lk_matched[i]
        assert nk == keep_g.sum()
        assert np.array_equal(out.lkf_t1_left[i, :nk], seeds[keep_g]) and np.array_equal(out.lkf_t1_right[i, :nk], out.lk_p1[i, :ns][keep_g])
        assert np.array_equal(out.lkf_t2_right[i, :nk], out.lk_p2[i, :ns][keep_g]) and np.array_equal(out.lkf_t2_left[i, :nk], out.lk_p3[i, :ns][keep_g])
        # K10: the ORB association filter (tracking.cpp:546-577)
        acc = orc.orb_match_filter(si, sd, ti, td, len(prev["dr"]), nl, prev["kl"], prev["kr"], 3.0)
        nm = out.n_orb_matched[i]
        assert nm == len(acc) and nm > 20
        cur_xy = np.stack([out.kps_left[i, :nl]["x"], out.kps_left[i, :nl]["y"]], 1)
        pl_xy = np.stack([prev["kl"]["x"], prev["kl"]["y"]], 1)
        pr_xy = np.stack([prev["kr"]["x"], prev["kr"]["y"]], 1)
        assert np.array_equal(out.orb_t1_left[i, :nm], pl_xy[acc[:, 0]]) and np.array_equal(out.orb_t1_right[i, :nm], pr_xy[acc[:, 1]])
        assert np.array_equal(out.orb_t2_left[i, :nm], cur_xy[acc[:, 2]])
    return _prev_of(out, i, lf, rf, ofast)


def _prev_of(out, i, lf, rf, ofast):
    nl, nr = out.n_left[i], out.n_right[i]
    return dict(l=lf, r=rf, dl=out.desc_left[i, :nl].copy(), dr=out.desc_right[i, :nr].copy(), fast=ofast,
                kl=out.kps_left[i, :nl].copy(), kr=out.kps_right[i, :nr].copy())


@pytest.mark.parametrize("w,h,nfeat,max_frames,n_total", [(640, 200, 500, 3, 7), (1241, 376, 1000, 4, 5)])
def test_frontend_sequence(pkg, orc, w, h, nfeat, max_frames, n_total):
    seed = 1000
    left, right = pkg.synth_sequence(seed, 13, n_total, w, h)  # frames 13..: crosses the canvas re-seed at frame 16
    ctx = pkg.Context(0, w, h, 2 * (max_frames + 1))
    ctx.orb_configure(nfeat, 1.2, 8, 20, 7)
    ctx.frontend_configure(w, h, pkg.MODE_ORB | pkg.MODE_LK, max_frames)
    kp_cap, pt_cap = ctx.orb_capacity(w, h), max(1024, w * h // 64)
    oracle = orc.OrbOracle(nfeat, 1.2, 8, 20, 7)
    prev = None  # oracle-side last frame
    f0 = 0
    while f0 < n_total:
        n = min(max_frames, n_total - f0)
        out = pkg.FrontendBuffers(n, kp_cap, pt_cap)
        ctx.frontend_batch(np.ascontiguousarray(left[f0:f0 + n]), np.ascontiguousarray(right[f0:f0 + n]), out)
        for i in range(n):
            prev = _check_unit(pkg, orc, oracle, out, i, left[f0 + i], right[f0 + i], prev)
        f0 += n
    ctx.close()


def test_bench_step_against_oracle(pkg, orc):
    """The batch bench.py times (config c1: 128 KITTI-sized frames of sequence seed 1000 from frame 0, nfeatures 1000) checked DIRECTLY
    against the oracle on a spread of its units -- the first, one either side of the canvas re-seed, the middle, the last -- each
    through every stage (ORB both eyes, both matches, FAST seeds, the 4-call LK chain, both filters)."""
    w, h, B, nfeat = 1241, 376, 128, 1000
    left, right = pkg.synth_sequence(1000, 0, B, w, h)
    ctx = pkg.Context(0, w, h, 2 * (B + 1))
    ctx.orb_configure(nfeat, 1.2, 8, 20, 7)
    ctx.frontend_configure(w, h, pkg.MODE_ORB | pkg.MODE_LK, B)
    kp_cap, pt_cap = ctx.orb_capacity(w, h), max(1024, w * h // 64)
    out = pkg.FrontendBuffers(B, kp_cap, pt_cap)
    ctx.frontend_batch(left, right, out)
    ctx.frontend_check()
    oracle = orc.OrbOracle(nfeat, 1.2, 8, 20, 7)
    for f in (0, 15, 16, 64, 127):
        prev = None if f == 0 else _prev_of(out, f - 1, left[f - 1], right[f - 1], orc.fast9(left[f - 1], 20))
        _check_unit(pkg, orc, oracle, out, f, left[f], right[f], prev)
    ctx.close()


def test_frontend_reset_and_partial_outputs(pkg):
    w, h = 640, 200
    left, right = pkg.synth_sequence(5, 0, 4, w, h)
    ctx = pkg.Context(0, w, h, 6)
    ctx.orb_configure(300, 1.2, 8, 20, 7)
    ctx.frontend_configure(w, h, pkg.MODE_ORB, 2)
    kp_cap = ctx.orb_capacity(w, h)
    a = pkg.FrontendBuffers(2, kp_cap, 1, fields=("n_left", "n_query", "stereo_idx"))
    ctx.frontend_batch(np.ascontiguousarray(left[:2]), np.ascontiguousarray(right[:2]), a)
    assert a.n_query[0] == 0 and a.n_query[1] == a.n_left[0]
    ctx.frontend_reset()
    b = pkg.FrontendBuffers(2, kp_cap, 1, fields=("n_left", "n_query", "stereo_idx"))
    ctx.frontend_batch(np.ascontiguousarray(left[:2]), np.ascontiguousarray(right[:2]), b)
    assert np.array_equal(a.n_left, b.n_left) and np.array_equal(a.n_query, b.n_query) and np.array_equal(a.stereo_idx, b.stereo_idx)
    ctx.close()


def test_frontend_full_hd_orb_with_matches(pkg, orc):
    """BASELINE.json configs[2]/[3] shape: ORB 2000 features at 1920x1080 plus both brute-force matches, batched."""
    w, h, nfeat = 1920, 1080, 2000
    left, right = pkg.synth_sequence(2000, 0, 3, w, h)
    ctx = pkg.Context(0, w, h, 6)
    ctx.orb_configure(nfeat, 1.2, 8, 20, 7)
    ctx.frontend_configure(w, h, pkg.MODE_ORB, 2)
    kp_cap = ctx.orb_capacity(w, h)
    oracle = orc.OrbOracle(nfeat, 1.2, 8, 20, 7)
    prev = None
    for f0, n in ((0, 2), (2, 1)):
        out = pkg.FrontendBuffers(n, kp_cap, 1)
        ctx.frontend_batch(np.ascontiguousarray(left[f0:f0 + n]), np.ascontiguousarray(right[f0:f0 + n]), out)
        for i in range(n):
            ok, od = oracle.extract(left[f0 + i])
            nl, nr = out.n_left[i], out.n_right[i]
            assert nl == len(ok) and np.array_equal(out.kps_left[i, :nl], ok)
            assert np.unpackbits(out.desc_left[i, :nl] ^ od).sum() <= 2
            if prev is not None:
                nq = out.n_query[i]
                si, sd = orc.match_hamming(prev[0], prev[1])
                ti, td = orc.match_hamming(prev[0], out.desc_left[i, :nl])
                assert np.array_equal(out.stereo_idx[i, :nq], si) and np.array_equal(out.stereo_dist[i, :nq], sd)
                assert np.array_equal(out.temporal_idx[i, :nq], ti) and np.array_equal(out.temporal_dist[i, :nq], td)
            prev = (out.desc_left[i, :nl].copy(), out.desc_right[i, :nr].copy())
    ctx.close()


def test_frontend_pipelined_equals_sequential(pkg):
    """Depth-2 ring: run(k+1) is issued before download(k); results must equal the synchronous svo_frontend_batch path."""
    w, h, B, nb = 640, 200, 3, 4
    left, right = pkg.synth_sequence(9, 0, B * nb, w, h)
    fields = ("kps_left", "desc_left", "n_left", "n_query", "stereo_idx", "temporal_idx", "fast_kps", "n_fast", "n_seed", "lk_p3", "lk_status")

    def make(max_images):
        ctx = pkg.Context(0, w, h, max_images)
        ctx.orb_configure(300, 1.2, 8, 20, 7)
        ctx.frontend_configure(w, h, pkg.MODE_ORB | pkg.MODE_LK, B)
        return ctx, ctx.orb_capacity(w, h), max(1024, w * h // 64)

    def views(k):
        return np.ascontiguousarray(left[k * B:(k + 1) * B]), np.ascontiguousarray(right[k * B:(k + 1) * B])

    seq, kc, pc = make(2 * (B + 1))
    ref = []
    for k in range(nb):
        o = pkg.FrontendBuffers(B, kc, pc, fields=fields)
        seq.frontend_batch(*views(k), o)
        ref.append(o)
    seq.close()

    pipe, kc, pc = make(2 * (2 * B + 1))
    got = [pkg.FrontendBuffers(B, kc, pc, fields=fields) for _ in range(nb)]
    pipe.frontend_upload(*views(0))
    pipe.frontend_run()
    for k in range(nb):
        if k + 1 < nb:
            pipe.frontend_upload(*views(k + 1))
            pipe.frontend_run()
        pipe.frontend_download(got[k])
    with pytest.raises(pkg.SvoError):
        pipe.frontend_download(got[0])  # nothing left to deliver
    pipe.close()
    for k in range(nb):
        for f in fields:
            a, b = getattr(ref[k], f), getattr(got[k], f)
            if f in ("n_left", "n_query", "n_fast", "n_seed"):
                assert np.array_equal(a, b), (k, f)
        for i in range(B):
            nl, nq, nf, ns = ref[k].n_left[i], ref[k].n_query[i], ref[k].n_fast[i], ref[k].n_seed[i]
            assert np.array_equal(ref[k].kps_left[i, :nl], got[k].kps_left[i, :nl]) and np.array_equal(ref[k].desc_left[i, :nl], got[k].desc_left[i, :nl])
            assert np.array_equal(ref[k].stereo_idx[i, :nq], got[k].stereo_idx[i, :nq]) and np.array_equal(ref[k].temporal_idx[i, :nq], got[k].temporal_idx[i, :nq])
            assert np.array_equal(ref[k].fast_kps[i, :nf], got[k].fast_kps[i, :nf])
            assert np.array_equal(ref[k].lk_p3[i, :ns], got[k].lk_p3[i, :ns]) and np.array_equal(ref[k].lk_status[i, :, :ns], got[k].lk_status[i, :, :ns])


KITTI_P1 = np.array([[718.856, 0, 607.1928, 0], [0, 718.856, 185.2157, 0], [0, 0, 1, 0]], np.float64)
KITTI_P2 = KITTI_P1.copy()
KITTI_P2[0, 3] = -386.1448


def test_triangulate_points(pkg, orc):
    gold = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "cv2_primitives.npz"))
    ctx = pkg.Context(0, 64, 64, 2)
    g4, g3 = ctx.triangulate_points(gold["tri_P1"], gold["tri_P2"], gold["tri_pts1"], gold["tri_pts2"])
    o4, o3 = orc.triangulate_points(gold["tri_P1"], gold["tri_P2"], gold["tri_pts1"], gold["tri_pts2"])
    ref3 = gold["tri_points3d"]  # cv2.triangulatePoints + convertPointsFromHomogeneous
    assert (np.abs(g3 - ref3) <= 2e-5 * np.abs(ref3) + 1e-6).all()
    assert (np.abs(g3 - o3) <= 2e-5 * np.abs(o3) + 1e-6).all()
    print("triangulation: bit-identical to cv2 on", int((g3 == ref3).all(1).sum()), "of", len(ref3), "points")
    assert ctx.triangulate_points(KITTI_P1, KITTI_P2, np.zeros((0, 2)), np.zeros((0, 2)))[1].shape == (0, 3)
    ctx.close()


def test_frontend_triangulated_lists(pkg, orc):
    """K12 inside the batch: 3-D points of the filtered ORB and LK lists == oracle triangulation of those lists."""
    w, h, B = 1241, 376, 3
    left, right = pkg.synth_sequence(1000, 2, B, w, h)
    ctx = pkg.Context(0, w, h, 2 * (B + 1))
    ctx.orb_configure(1000, 1.2, 8, 20, 7)
    ctx.frontend_configure(w, h, pkg.MODE_ORB | pkg.MODE_LK, B)
    ctx.frontend_set_projection(KITTI_P1, KITTI_P2)
    out = pkg.FrontendBuffers(B, ctx.orb_capacity(w, h), max(1024, w * h // 64))
    ctx.frontend_batch(left, right, out)
    for i in range(1, B):
        for n, a, b, xyz in ((out.n_orb_matched[i], out.orb_t1_left[i], out.orb_t1_right[i], out.orb_points3d[i]),
                             (out.n_lk_matched[i], out.lkf_t1_left[i], out.lkf_t1_right[i], out.lk_points3d[i])):
            assert n > 50
            _, o3 = orc.triangulate_points(KITTI_P1, KITTI_P2, a[:n], b[:n])
            assert (np.abs(xyz[:n] - o3) <= 2e-5 * np.abs(o3) + 1e-6).all()
        # synthetic disparity is 8 px: depth = fx * baseline / 8 for the bulk of the LK tracks
        z = out.lk_points3d[i, :out.n_lk_matched[i], 2]
        assert abs(np.median(z) - 386.1448 / 8) < 1.0
    ctx.close()


def test_frontend_equals_reference_tracking_source(pkg, orc):
    """The batch front-end against tracking.cpp:534-660 compiled from the reference's own text (oracle/_ref/libtracking_ref.so):
    the ORB association lists bit-exact (same keypoints / descriptors in), the surviving LK tracks identical in number and
    within 0.01 px."""
    if orc.ref_tracking_lib() is None:
        pytest.skip("oracle/_ref/libtracking_ref.so not present (built where /root/reference exists)")
    w, h, B = 1241, 376, 4
    left, right = pkg.synth_sequence(1000, 14, B, w, h)  # crosses the canvas re-seed at frame 16
    ctx = pkg.Context(0, w, h, 2 * (B + 1))
    ctx.orb_configure(1000, 1.2, 8, 20, 7)
    ctx.frontend_configure(w, h, pkg.MODE_ORB | pkg.MODE_LK, B)
    out = pkg.FrontendBuffers(B, ctx.orb_capacity(w, h), max(1024, w * h // 64))
    ctx.frontend_batch(left, right, out)
    for i in range(1, B):
        nl0, nr0, nl1 = out.n_left[i - 1], out.n_right[i - 1], out.n_left[i]
        t2l, t1l, t1r = orc.ref_orb_robust_match(out.kps_left[i - 1, :nl0], out.desc_left[i - 1, :nl0], out.kps_right[i - 1, :nr0],
                                                 out.desc_right[i - 1, :nr0], out.kps_left[i, :nl1], out.desc_left[i, :nl1], 3.0)
        nm = out.n_orb_matched[i]
        assert nm == len(t2l) and nm > 20
        assert np.array_equal(out.orb_t2_left[i, :nm], t2l) and np.array_equal(out.orb_t1_left[i, :nm], t1l)
        assert np.array_equal(out.orb_t1_right[i, :nm], t1r)
        ns = out.n_seed[i]
        seeds = np.stack([out.fast_kps[i - 1, :ns]["x"], out.fast_kps[i - 1, :ns]["y"]], 1)
        r_t2l, r_t2r, r_t1l, r_t1r = orc.ref_lk_robust_track(left[i - 1], right[i - 1], left[i], right[i], seeds, 3.0)
        nk = out.n_lk_matched[i]
        assert nk == len(r_t1l) and nk > 50, (nk, len(r_t1l))
        assert np.array_equal(out.lkf_t1_left[i, :nk], r_t1l)
        for g, r in ((out.lkf_t1_right, r_t1r), (out.lkf_t2_right, r_t2r), (out.lkf_t2_left, r_t2l)):
            assert np.abs(g[i, :nk] - r).max() <= 0.01
    ctx.close()


def test_frontend_overlap_equals_single_stream(pkg):
    """The ORB tail on the side stream (default) gives exactly the outputs of the single-stream schedule."""
    w, h, B = 640, 200, 4
    left, right = pkg.synth_sequence(31, 0, 2 * B, w, h)
    res = []
    for overlap in (True, False):
        ctx = pkg.Context(0, w, h, 2 * (B + 1))
        ctx.orb_configure(400, 1.2, 8, 20, 7)
        ctx.frontend_configure(w, h, pkg.MODE_ORB | pkg.MODE_LK, B)
        ctx.frontend_set_overlap(overlap)
        outs = []
        for b in range(2):
            out = pkg.FrontendBuffers(B, ctx.orb_capacity(w, h), max(1024, w * h // 64))
            ctx.frontend_batch(np.ascontiguousarray(left[b * B:(b + 1) * B]), np.ascontiguousarray(right[b * B:(b + 1) * B]), out)
            outs.append(out)
        res.append(outs)
        ctx.close()
    for a, b in zip(*res):
        for name in pkg.FrontendBuffers.ALL:
            x, y = getattr(a, name), getattr(b, name)
            if x is not None and name not in ("orb_points3d", "lk_points3d"):
                assert np.array_equal(x, y), name


def test_frontend_lk_full_hd_many_points(pkg, orc):
    """Full HD, LK mode only: several thousand FAST corners per frame through the fused 4-call chain, against the oracle chain."""
    w, h, B = 1920, 1080, 2
    left, right = pkg.synth_sequence(2000, 3, B, w, h)
    ctx = pkg.Context(0, w, h, 2 * (B + 1))
    ctx.frontend_configure(w, h, pkg.MODE_LK, B)
    pt_cap = max(1024, w * h // 64)
    out = pkg.FrontendBuffers(B, 64, pt_cap, fields=("fast_kps", "n_fast", "n_seed", "lk_p1", "lk_p2", "lk_p3", "lk_p0r", "lk_status", "n_lk_matched"))
    ctx.frontend_batch(left, right, out)
    ns = int(out.n_seed[1])
    assert ns > 2048, ns
    seeds = np.stack([out.fast_kps[0, :ns]["x"], out.fast_kps[0, :ns]["y"]], 1)
    (p1, p2, p3, p0r), sts = _oracle_chain(orc, left[0], right[0], left[1], right[1], seeds)
    cum = np.ones(ns, bool)
    for k, (gp, op) in enumerate(zip((out.lk_p1, out.lk_p2, out.lk_p3, out.lk_p0r), (p1, p2, p3, p0r))):
        assert np.array_equal(out.lk_status[1, k, :ns][cum], sts[k][cum])
        cum &= sts[k] == 1
        assert np.abs(gp[1, :ns] - op)[cum].max() <= 0.01
    assert cum.sum() > 1000
    ctx.close()
