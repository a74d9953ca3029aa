"""GPU: the state machine of the throughput mode -- capacity overflows are never silent, a download that fails on a caller
capacity can be retried, standalone calls and batches may be mixed on one context, and the LK chain's `err` output equals the
oracle's (the reference passes `err` to every calcOpticalFlowPyrLK call, tracking.cpp:593-618)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _dots(n, h, w, step=4):
    """Isolated bright dots on a dark field: every dot is a FAST corner that survives suppression (1 / step^2 of the pixels)."""
    img = np.full((n, h, w), 10, np.uint8)
    img[:, 4:h - 4:step, 4:w - 4:step] = 250
    return img


def test_overflow_is_reported_on_download_and_on_check(pkg):
    w, h, B = 640, 200, 2
    ctx = pkg.Context(0, w, h, 2 * (2 * B + 1))
    ctx.frontend_configure(w, h, pkg.MODE_LK, B)
    pt_cap = max(1024, w * h // 64)
    dots = _dots(B, h, w)  # ~ w * h / 16 corners per frame: four times the device's pt_cap
    good_l, good_r = pkg.synth_sequence(3, 0, B, w, h)
    out = pkg.FrontendBuffers(B, 1, pt_cap, fields=("n_fast", "n_seed", "fast_kps"))
    with pytest.raises(pkg.SvoError) as e:
        ctx.frontend_batch(dots, dots, out)
    assert e.value.code == pkg.SVO_ERR_CAPACITY and "device capacity" in str(e.value)
    # the flags are per run: a following good batch is delivered, and the overflowed run is gone from the queue
    ctx.frontend_reset()
    ctx.frontend_batch(good_l, good_r, out)
    assert out.n_fast.min() > 50
    # device-resident pipeline: the run is retired without a download; svo_frontend_check() must still see the overflow
    ctx.frontend_reset()
    ctx.frontend_upload(dots, dots)
    ctx.frontend_run()
    ctx.frontend_download(None)
    ctx.frontend_upload(good_l, good_r)  # later good runs do not hide it
    ctx.frontend_run()
    ctx.frontend_download(None)
    with pytest.raises(pkg.SvoError) as e:
        ctx.frontend_check()
    assert e.value.code == pkg.SVO_ERR_CAPACITY
    ctx.frontend_check()  # reported once, then clear
    ctx.close()


def test_download_with_small_caller_buffers_can_be_retried(pkg):
    w, h, B = 640, 200, 2
    left, right = pkg.synth_sequence(5, 0, 2 * B, w, h)
    ctx = pkg.Context(0, w, h, 2 * (2 * B + 1))
    ctx.orb_configure(300, 1.2, 8, 20, 7)
    ctx.frontend_configure(w, h, pkg.MODE_ORB | pkg.MODE_LK, B)
    kp_cap, pt_cap = ctx.orb_capacity(w, h), max(1024, w * h // 64)
    ref = pkg.FrontendBuffers(B, kp_cap, pt_cap)
    ctx.frontend_batch(np.ascontiguousarray(left[:B]), np.ascontiguousarray(right[:B]), ref)
    ctx.frontend_reset()
    ctx.frontend_upload(np.ascontiguousarray(left[:B]), np.ascontiguousarray(right[:B]))
    ctx.frontend_run()
    for bad in (pkg.FrontendBuffers(B, kp_cap - 1, pt_cap), pkg.FrontendBuffers(B, kp_cap, 8)):
        with pytest.raises(pkg.SvoError) as e:
            ctx.frontend_download(bad)
        assert e.value.code == pkg.SVO_ERR_CAPACITY and "kept queued" in str(e.value)
        assert bad.n_left.sum() == 0 and bad.kps_left["x"].sum() == 0  # nothing was written
    again = pkg.FrontendBuffers(B, kp_cap, pt_cap)
    ctx.frontend_download(again)  # the same run, delivered now
    for name in ("n_left", "n_query", "n_fast", "n_seed", "n_lk_matched"):
        assert np.array_equal(getattr(ref, name), getattr(again, name)), name
    for i in range(B):  # entries beyond the per-frame counts are unspecified
        nl, nq, nf, ns = int(ref.n_left[i]), int(ref.n_query[i]), int(ref.n_fast[i]), int(ref.n_seed[i])
        assert np.array_equal(ref.kps_left[i, :nl], again.kps_left[i, :nl]) and np.array_equal(ref.desc_left[i, :nl], again.desc_left[i, :nl])
        assert np.array_equal(ref.stereo_idx[i, :nq], again.stereo_idx[i, :nq]) and np.array_equal(ref.fast_kps[i, :nf], again.fast_kps[i, :nf])
        assert np.array_equal(ref.lk_p1[i, :ns], again.lk_p1[i, :ns]) and np.array_equal(ref.lk_status[i, :, :ns], again.lk_status[i, :, :ns])
        assert np.array_equal(ref.lk_err[i, :, :ns], again.lk_err[i, :, :ns])
    with pytest.raises(pkg.SvoError):
        ctx.frontend_download(again)  # nothing left
    ctx.frontend_upload(np.ascontiguousarray(left[B:]), np.ascontiguousarray(right[B:]))
    ctx.frontend_run()
    ctx.frontend_discard()
    with pytest.raises(pkg.SvoError):
        ctx.frontend_download(again)
    ctx.close()


def test_standalone_calls_between_batches(pkg, orc):
    """svo_fast_detect with a cap above the front-end's pt_cap re-allocates the point buffers; the context must not
    keep using the stale throughput-mode layout, and a standalone call must not leave a clobbered `last frame` behind."""
    w, h, B = 640, 200, 2
    left, right = pkg.synth_sequence(9, 0, 2 * B, w, h)
    ctx = pkg.Context(0, w, h, 2 * (B + 1))
    ctx.orb_configure(300, 1.2, 8, 20, 7)
    ctx.frontend_configure(w, h, pkg.MODE_ORB | pkg.MODE_LK, B)
    kp_cap, pt_cap = ctx.orb_capacity(w, h), max(1024, w * h // 64)
    a = pkg.FrontendBuffers(B, kp_cap, pt_cap)
    ctx.frontend_batch(np.ascontiguousarray(left[:B]), np.ascontiguousarray(right[:B]), a)
    # a standalone call with the SAME capacity: allowed, drops the carried frame
    k = ctx.fast_detect(left[0], 20, True, cap=pt_cap)
    assert np.array_equal(k, orc.fast9(left[0], 20))
    b = pkg.FrontendBuffers(B, kp_cap, pt_cap)
    ctx.frontend_batch(np.ascontiguousarray(left[B:]), np.ascontiguousarray(right[B:]), b)
    assert b.n_query[0] == 0 and b.n_seed[0] == 0 and b.n_query[1] == b.n_left[0]  # like after a reset
    # a standalone call that grows the point buffers: the throughput mode must be configured again
    k = ctx.fast_detect(left[1], 20, True, cap=w * h // 16)
    assert np.array_equal(k, orc.fast9(left[1], 20))
    with pytest.raises(pkg.SvoError):
        ctx.frontend_upload(np.ascontiguousarray(left[:B]), np.ascontiguousarray(right[:B]))
    ctx.frontend_configure(w, h, pkg.MODE_ORB | pkg.MODE_LK, B)
    c = pkg.FrontendBuffers(B, kp_cap, max(1024, w * h // 16))
    ctx.frontend_batch(np.ascontiguousarray(left[:B]), np.ascontiguousarray(right[:B]), c)
    for name in ("n_left", "n_query", "n_fast", "n_lk_matched", "n_seed"):
        assert np.array_equal(getattr(a, name), getattr(c, name)), name
    for i in range(B):  # entries beyond the per-frame counts are unspecified
        nl, nq = int(a.n_left[i]), int(a.n_query[i])
        assert np.array_equal(a.kps_left[i, :nl], c.kps_left[i, :nl]) and np.array_equal(a.desc_left[i, :nl], c.desc_left[i, :nl])
        assert np.array_equal(a.stereo_idx[i, :nq], c.stereo_idx[i, :nq])
    ns = int(a.n_seed[1])
    assert np.array_equal(a.lk_p1[1, :ns], c.lk_p1[1, :ns]) and np.array_equal(a.lk_status[1, :, :ns], c.lk_status[1, :, :ns])
    # undelivered runs block standalone calls instead of being clobbered
    ctx.frontend_upload(np.ascontiguousarray(left[B:]), np.ascontiguousarray(right[B:]))
    ctx.frontend_run()
    with pytest.raises(pkg.SvoError):
        ctx.orb_extract(left[0])
    ctx.frontend_download(c)
    ctx.orb_extract(left[0])
    ctx.close()


def test_lk_chain_err_equals_oracle(pkg, orc):
    w, h, B = 640, 200, 3
    left, right = pkg.synth_sequence(1000, 14, B, w, h)
    ctx = pkg.Context(0, w, h, 2 * (B + 1))
    ctx.frontend_configure(w, h, pkg.MODE_LK, B)
    pt_cap = max(1024, w * h // 64)
    out = pkg.FrontendBuffers(B, 1, pt_cap)
    ctx.frontend_batch(left, right, out)
    checked = 0
    for f in range(1, B):
        seeds = orc.fast9(left[f - 1], 20)
        pts = np.stack([seeds["x"], seeds["y"]], 1)
        ns = int(out.n_seed[f])
        assert ns == len(pts)
        chain = ((left[f - 1], right[f - 1]), (right[f - 1], right[f]), (right[f], left[f]), (left[f], left[f - 1]))
        cum = np.ones(ns, bool)
        src = pts
        for k, (a, b) in enumerate(chain):
            p, s, e = orc.lk_track(a, b, src)
            ok = cum & (s == 1)
            ge = out.lk_err[f, k, :ns]
            assert np.abs(ge[ok] - e[ok]).max() <= 0.02, f"err of call {k + 1}"  # mean absolute window difference / 32
            assert np.all(ge[cum & (s == 0)] == 0)
            cum = ok
            src = p
            checked += int(ok.sum())
    assert checked > 500
    # switched off: the field is left alone
    ctx.frontend_set_lk_err(False)
    ctx.frontend_reset()
    out2 = pkg.FrontendBuffers(B, 1, pt_cap)
    ctx.frontend_batch(left, right, out2)
    assert out2.lk_err.sum() == 0
    for f in range(1, B):  # entries beyond the per-frame counts are unspecified
        ns = int(out.n_seed[f])
        assert np.array_equal(out2.lk_p1[f, :ns], out.lk_p1[f, :ns]) and np.array_equal(out2.lk_status[f, :, :ns], out.lk_status[f, :, :ns])
    ctx.close()


def test_match_4000_by_4000(pkg, orc):
    """BASELINE.json configs[3]: both brute-force matches at 4000 features."""
    rng = np.random.default_rng(4000)
    t = rng.integers(0, 256, (4000, 32), dtype=np.uint8)
    q = t[rng.permutation(4000)].copy()
    flip = rng.integers(0, 256, (4000, 32), dtype=np.uint8) & rng.integers(0, 256, (4000, 32), dtype=np.uint8) & rng.integers(0, 256, (4000, 32), dtype=np.uint8)
    q ^= flip  # ~12 % of the bits flipped: distances well below random pairs, many near-ties among the 4000 candidates
    q[:50] = t[7]  # 50 queries with several exact minima: first-index tie-break
    t[100:110] = t[7]
    ctx = pkg.Context(0, 64, 64, 2)
    gi, gd = ctx.match_hamming(q, t)
    oi, od = orc.match_hamming(q, t)
    assert np.array_equal(gi, oi) and np.array_equal(gd, od)
    assert np.all(gi[:50] == 7)
    ctx.close()


def test_right_image_at_4k(pkg, orc):
    """The right image of the 3840x2160 pair (tests/test_orb_gpu.py checks the left one against the oracle)."""
    w, h, nfeat = 3840, 2160, 4000
    left, right = pkg.synth_stereo_frame(3000, 0, w, h)
    ctx = pkg.Context(0, w, h, 2)
    ctx.orb_configure(nfeat, 1.2, 8, 20, 7)
    kps, desc = ctx.orb_extract_batch([left, right])
    o = orc.OrbOracle(nfeat, 1.2, 8, 20, 7)
    ok, od = o.extract(right)
    assert len(kps[1]) == len(ok) and np.array_equal(kps[1], ok)
    assert int(np.unpackbits(desc[1] ^ od).sum()) <= 10
    for l in (0, 3, 7):
        assert np.array_equal(ctx.orb_read_level(1, l), o.level(l))
    ctx.close()


def test_frame_source_feeds_the_frontend(pkg, tmp_path):
    """System::NextFrame_kitti replacement: PNG files on disk -> decoder threads -> pinned ring -> svo_frontend_upload; the results
    equal those of the same frames handed over from memory."""
    w, h, B, F = 640, 200, 4, 10
    left, right = pkg.synth_sequence(21, 0, F, w, h)
    pkg.write_sequence(str(tmp_path), left, right, fmt="png")
    ctx = pkg.Context(0, w, h, 2 * (2 * B + 1))
    ctx.orb_configure(300, 1.2, 8, 20, 7)
    ctx.frontend_configure(w, h, pkg.MODE_ORB | pkg.MODE_LK, B)
    kp_cap, pt_cap = ctx.orb_capacity(w, h), max(1024, w * h // 64)
    ref = []
    for f0 in range(0, F, B):
        n = min(B, F - f0)
        o = pkg.FrontendBuffers(n, kp_cap, pt_cap)
        ctx.frontend_batch(np.ascontiguousarray(left[f0:f0 + n]), np.ascontiguousarray(right[f0:f0 + n]), o)
        ref.append(o)
    ctx.frontend_reset()
    src = pkg.FrameSource(str(tmp_path), w, h, B, ring_batches=3, threads=3)
    assert src.is_pinned() and src.frame_count() == F
    i = 0
    while True:
        l, r, n = src.next_ptr()
        if n == 0:
            break
        o = pkg.FrontendBuffers(n, kp_cap, pt_cap)
        ctx.frontend_upload_ptr(l, r, n, w)
        ctx.frontend_run()
        ctx.frontend_download(o)
        q = ref[i]
        for name in ("n_left", "n_right", "n_query", "n_fast", "n_seed", "n_lk_matched", "n_orb_matched"):
            assert np.array_equal(getattr(o, name), getattr(q, name)), (i, name)
        for j in range(n):  # entries beyond the per-frame counts are unspecified
            nl, nr, nq, nf, ns = int(q.n_left[j]), int(q.n_right[j]), int(q.n_query[j]), int(q.n_fast[j]), int(q.n_seed[j])
            assert np.array_equal(o.kps_left[j, :nl], q.kps_left[j, :nl]) and np.array_equal(o.desc_right[j, :nr], q.desc_right[j, :nr])
            assert np.array_equal(o.stereo_idx[j, :nq], q.stereo_idx[j, :nq]) and np.array_equal(o.temporal_idx[j, :nq], q.temporal_idx[j, :nq])
            assert np.array_equal(o.fast_kps[j, :nf], q.fast_kps[j, :nf]) and np.array_equal(o.lk_p3[j, :ns], q.lk_p3[j, :ns])
            assert np.array_equal(o.lk_status[j, :, :ns], q.lk_status[j, :, :ns])
        i += 1
    assert i == len(ref)
    src.close()
    ctx.close()


def test_standalone_device_filters_equal_oracle(pkg, orc):
    """svo_orb_match_filter / svo_lk_chain_filter (K10 / K11 on caller arrays; what the drop-in Tracking class calls) against the
    oracle's restatements of tracking.cpp:546-577 and :623-660."""
    w, h = 640, 200
    l0, r0 = pkg.synth_stereo_frame(55, 2, w, h)
    l1, r1 = pkg.synth_stereo_frame(55, 3, w, h)
    ctx = pkg.Context(0, w, h, 2)
    ctx.orb_configure(500, 1.2, 8, 20, 7)
    ks, ds = ctx.orb_extract_batch([l0, r0])
    kl0, kr0, dl0, dr0 = ks[0], ks[1], ds[0], ds[1]
    ks, ds = ctx.orb_extract_batch([l1, r1])
    kl1, dl1 = ks[0], ds[0]
    i1, d1 = ctx.match_hamming(dl0, dr0)
    i2, d2 = ctx.match_hamming(dl0, dl1)
    for fme in (3.0, 0.5):
        acc = orc.orb_match_filter(i1, d1, i2, d2, len(kr0), len(kl1), kl0, kr0, fme)
        t2, t1l, t1r = ctx.orb_match_filter(kl0, kr0, kl1, i1, d1, i2, d2, fme)
        assert len(t2) == len(acc) and len(acc) > (20 if fme > 1 else 0)
        xy = lambda k: np.stack([k["x"], k["y"]], 1)
        assert np.array_equal(t1l, xy(kl0)[acc[:, 0]]) and np.array_equal(t1r, xy(kr0)[acc[:, 1]]) and np.array_equal(t2, xy(kl1)[acc[:, 2]])
    # quirk Q3: fewer current keypoints than queries -> only the first des_index queries are considered
    m = 40
    t2, _, _ = ctx.orb_match_filter(kl0, kr0, kl1[:m], i1, d1, np.minimum(i2, m - 1), d2, 3.0)
    acc = orc.orb_match_filter(i1, d1, np.minimum(i2, m - 1), d2, len(kr0), m, kl0, kr0, 3.0)
    assert len(t2) == len(acc) and (len(acc) == 0 or acc[:, 0].max() < m)
    # LK chain filter
    seeds = orc.fast9(l0, 20)
    p0 = np.stack([seeds["x"], seeds["y"]], 1)
    p1, s1, _ = ctx.lk_track(l0, r0, p0)
    p2, s2, _ = ctx.lk_track(r0, r1, p1)
    p3, s3, _ = ctx.lk_track(r1, l1, p2)
    p0r, s4, _ = ctx.lk_track(l1, l0, p3)
    p2[5] = (-1.0, 4.0)  # a negative coordinate drops the track
    keep = orc.lk_chain_filter(p0, p1, p2, p3, p0r, s1, s2, s3, s4, 3.0)
    out = ctx.lk_chain_filter(p0, p1, p2, p3, p0r, s1, s2, s3, s4, 3.0)
    assert len(out[0]) == keep.sum() and 0 < keep.sum() < len(p0) and not keep[5]
    for g, o in zip(out, (p0, p1, p2, p3, p0r)):
        assert np.array_equal(g, o[keep])
    ctx.close()


def test_no_kernel_writes_outside_its_buffers():
    """SVO_GUARD=1: every device buffer of the library sits between two 4 KB guard bands (svo_dev_malloc); after the whole C ABI has
    run at awkward sizes (tests/run_guarded.py) no band byte may have changed.  The write-overrun check of this suite (the pool's
    compute-sanitizer is closed)."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, SVO_GUARD="1")
    r = subprocess.run([sys.executable, os.path.join(here, "run_guarded.py")], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    last = r.stdout.strip().splitlines()[-1].split()
    assert last[0] == "GUARD" and int(last[1]) == 0 and int(last[2]) > 50
