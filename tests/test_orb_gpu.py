"""GPU parity: ORBextractor::operator() through the C ABI against the CPU oracle (bit-exact for integer stages)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CONFIGS = [
    # (w, h, nfeatures, seed)           C1-shaped, small, odd-sized, and a larger one
    (1241, 376, 1000, 1000),
    (640, 200, 500, 7),
    (333, 257, 300, 11),
    (1920, 1080, 2000, 2000),   # BASELINE.json configs[2]
    (3840, 2160, 4000, 3000),   # BASELINE.json configs[3]
]


def _extract_both(pkg, orc, w, h, nfeat, seed, frame=0):
    left, right = pkg.synth_stereo_frame(seed, frame, w, h)
    ctx = pkg.Context(0, w, h, 2)
    ctx.orb_configure(nfeat, 1.2, 8, 20, 7)
    kps, desc = ctx.orb_extract_batch([left, right])
    oracles = []
    for img in (left, right):
        o = orc.OrbOracle(nfeat, 1.2, 8, 20, 7)
        ok, od = o.extract(img)
        oracles.append((o, ok, od))
    return ctx, kps, desc, oracles


@pytest.mark.parametrize("w,h,nfeat,seed", CONFIGS)
def test_pyramid_candidates_keypoints(pkg, orc, w, h, nfeat, seed):
    ctx, kps, desc, oracles = _extract_both(pkg, orc, w, h, nfeat, seed)
    total_bits = 0
    for slot, (o, ok, od) in enumerate(oracles):
        # K1: pyramid pixels bit-exact
        for l in range(8):
            assert np.array_equal(ctx.orb_read_level(slot, l), o.level(l)), f"pyramid level {l} slot {slot}"
        # K2: candidates of the cell loop, bit-exact including order
        for l in range(8):
            g = ctx.orb_read_candidates(slot, l)
            c = o.candidates(l)
            assert len(g) == len(c), f"candidate count level {l}: {len(g)} vs {len(c)}"
            for f in ("x", "y", "response"):
                assert np.array_equal(g[f], c[f]), f"candidates {f} level {l}"
        # K3 + K4 + epilogue: selected keypoints bit-exact including order, angle bit-exact
        k = kps[slot]
        assert len(k) == len(ok) and len(k) >= min(nfeat, 50)
        for f in ("x", "y", "size", "response", "octave", "class_id", "angle"):
            assert np.array_equal(k[f], ok[f]), f"keypoint field {f} slot {slot}"
        # K5: descriptor bits (identical angles here, so any difference would be cos/sin rounding)
        bits = int(np.unpackbits(desc[slot] ^ od).sum())
        total_bits += bits
        assert bits <= max(2, len(k) * 256 // 100000), f"descriptor bit mismatches {bits} of {len(k) * 256}"
        if w * h > 4_000_000 and slot == 0:
            break  # one 4K image is enough for the oracle's run time
    print(f"{w}x{h}: descriptor bit mismatch rate {total_bits}/{sum(len(k) for k in kps) * 256}")
    ctx.close()


def test_batch_equals_single(pkg):
    w, h = 640, 200
    imgs = [pkg.synth_stereo_frame(3, f, w, h)[0] for f in range(5)]
    ctx = pkg.Context(0, w, h, 8)
    ctx.orb_configure(500, 1.2, 8, 20, 7)
    kb, db = ctx.orb_extract_batch(imgs)
    for i, img in enumerate(imgs):
        k1, d1 = ctx.orb_extract(img)
        assert np.array_equal(k1, kb[i]) and np.array_equal(d1, db[i])
    ctx.close()


def test_flat_and_empty_images(pkg):
    ctx = pkg.Context(0, 640, 480, 1)
    ctx.orb_configure(500, 1.2, 8, 20, 7)
    k, d = ctx.orb_extract(np.full((480, 640), 77, np.uint8))  # flat image -> 0 keypoints (ORBextractor.cpp:1013-1014)
    assert len(k) == 0 and d.shape == (0, 32)
    k, d = ctx.orb_extract(np.zeros((0, 0), np.uint8))  # empty image -> silent return (ORBextractor.cpp:993-994)
    assert len(k) == 0
    ctx.close()


def test_strided_input(pkg, orc):
    w, h = 500, 300
    big = np.zeros((h, w + 37), np.uint8)
    left, _ = pkg.synth_stereo_frame(5, 0, w, h)
    big[:, :w] = left
    ctx = pkg.Context(0, w, h, 1)
    ctx.orb_configure(400, 1.2, 8, 20, 7)
    import ctypes as C
    cap = ctx.orb_capacity(w, h)
    kps = np.zeros(cap, pkg.KP_DTYPE)
    desc = np.zeros((cap, 32), np.uint8)
    n = C.c_int()
    rc = pkg.lib().svo_orb_extract(ctx._h, big.ctypes.data_as(C.c_void_p), w, h, C.c_size_t(big.strides[0]),
                                   kps.ctypes.data_as(C.c_void_p), desc.ctypes.data_as(C.c_void_p), cap, C.byref(n))
    assert rc == 0
    k2, d2 = ctx.orb_extract(left)
    assert np.array_equal(kps[:n.value], k2) and np.array_equal(desc[:n.value], d2)
    ctx.close()


@pytest.mark.parametrize("nq,nt", [(1000, 1000), (1, 1), (130, 7), (257, 4000), (5, 0)])
def test_match_hamming(pkg, orc, nq, nt):
    rng = np.random.default_rng(nq * 7 + nt)
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    if nt > 10:  # force ties and exact matches: first-minimum tie-break must hold
        t[7] = t[3]
        q[0] = t[3]
        q[nq // 2] = t[min(nt - 1, 9)]
        t[min(nt - 1, 9)] = t[1]
    ctx = pkg.Context(0, 64, 64, 1)
    gi, gd = ctx.match_hamming(q, t)
    if nt == 0:
        assert (gi == -1).all()
    else:
        oi, od = orc.match_hamming(q, t)
        assert np.array_equal(gi, oi) and np.array_equal(gd, od)
    ctx.close()


def test_match_real_descriptors(pkg, orc):
    w, h = 1241, 376
    left, right = pkg.synth_stereo_frame(1000, 1, w, h)
    ctx = pkg.Context(0, w, h, 2)
    ctx.orb_configure(1000, 1.2, 8, 20, 7)
    (kl, kr), (dl, dr) = ctx.orb_extract_batch([left, right])
    gi, gd = ctx.match_hamming(dl, dr)
    oi, od = orc.match_hamming(dl, dr)
    assert np.array_equal(gi, oi) and np.array_equal(gd, od)
    # synthetic disparity is 8 px: most good matches are on the same row, 8 px apart
    good = gd <= 30
    dx = kl["x"][good] - kr["x"][gi[good]]
    assert good.sum() > 300 and np.median(dx) == pytest.approx(8.0, abs=1.0)
    ctx.close()


def test_cuda_equals_reference_source_vectors(pkg, orc):
    """The CUDA path against outputs of the REFERENCE'S OWN ORBextractor.cpp (compiled against oracle/ref_shim, monotone
    allocator; tests/golden/ref_orbextractor.npz made by oracle/make_golden_ref.py) and, where oracle/_ref is present, against
    that library live: keypoints bit-exact in order with all fields, descriptors within 2 bits per image."""
    import ast
    import os
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_orbextractor.npz"))
    live = orc.ref_lib() is not None
    for case in gold["cases"]:
        key, seed, frame, w, h, nf, sf, nl, ini, mn = ast.literal_eval(str(case))
        img, _ = pkg.synth_stereo_frame(seed, frame, w, h)
        ctx = pkg.Context(0, w, h, 1)
        ctx.orb_configure(nf, sf, nl, ini, mn)
        k, d = ctx.orb_extract(img)
        refs = [(gold[key + "_kps"], gold[key + "_desc"])]
        if live:
            refs.append(orc.RefOrb(nf, sf, nl, ini, mn, monotone=True).extract(img))
        for rk, rd in refs:
            assert len(k) == len(rk)
            for f in ("x", "y", "size", "angle", "response", "octave", "class_id"):
                assert np.array_equal(k[f], rk[f]), (key, f)
            assert int(np.unpackbits(d ^ rd).sum()) <= 2, key
        ctx.close()
    print("reference library live:", live)
