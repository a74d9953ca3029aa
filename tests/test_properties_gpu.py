"""GPU: size-independent properties and edge cases -- random geometries vs the oracle, shard equivalence (the multi-GPU
rule of SURVEY.md 8e checked on one device), error paths of the C ABI."""
import ctypes as C
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_random_geometries_match_oracle(pkg, orc):
    """Odd widths/heights, different level counts and scale factors exercise the cell grid, tile split and octree roots."""
    rng = np.random.default_rng(12345)
    cases = [(int(rng.integers(120, 900)), int(rng.integers(100, 500)), int(rng.integers(100, 1500)), int(rng.integers(1, 9)),
              float(rng.choice([1.2, 1.1, 1.5, 1.33]))) for _ in range(8)]
    cases += [(97, 97, 50, 2, 1.2), (1241, 376, 1000, 8, 2.0), (640, 480, 700, 5, 1.2), (400, 1000, 300, 4, 1.2)]
    for w, h, nfeat, nlev, sf in cases:
        if w < h / 2 + 1:  # nIni == 0 is undefined in the reference
            continue
        left, _ = pkg.synth_stereo_frame(w * 7 + h, 1, w, h)
        ctx = pkg.Context(0, w, h, 1)
        ctx.orb_configure(nfeat, sf, nlev, 20, 7)
        try:
            k, d = ctx.orb_extract(left)
        except pkg.SvoError as e:  # levels too small for a cell grid are rejected loudly, never mis-computed
            assert e.code == pkg.SVO_ERR_INVALID, (w, h, nlev, sf, str(e))
            ctx.close()
            continue
        o = orc.OrbOracle(nfeat, sf, nlev, 20, 7)
        ok, od = o.extract(left)
        assert len(k) == len(ok), (w, h, nfeat, nlev, sf, len(k), len(ok))
        assert np.array_equal(k, ok), (w, h, nfeat, nlev, sf)
        assert np.unpackbits(d ^ od).sum() <= 2
        for l in range(nlev):
            assert np.array_equal(ctx.orb_read_level(0, l), o.level(l)), (w, h, l, sf)
        ctx.close()


@pytest.mark.parametrize("ini_th,min_th", [(20, 7), (30, 10), (12, 5), (40, 15), (9, 3), (7, 7), (100, 31)])
def test_fast_threshold_pairs_match_oracle(pkg, orc, ini_th, min_th):
    """iniThFAST / minThFAST other than the reference's defaults: minTh + 1 a power of two (OR-combined pre-test) or not
    (per-difference pre-test), equal thresholds, and a high pair that leaves most cells empty."""
    w, h = 752, 480
    left, _ = pkg.synth_stereo_frame(4242, 3, w, h)
    ctx = pkg.Context(0, w, h, 1)
    ctx.orb_configure(800, 1.2, 6, ini_th, min_th)
    k, d = ctx.orb_extract(left)
    o = orc.OrbOracle(800, 1.2, 6, ini_th, min_th)
    ok, od = o.extract(left)
    for l in range(6):
        g, c = ctx.orb_read_candidates(0, l), o.candidates(l)
        assert len(g) == len(c), (l, len(g), len(c))
        for f in ("x", "y", "response"):
            assert np.array_equal(g[f], c[f]), (l, f)
    assert len(k) == len(ok) and np.array_equal(k, ok)
    assert np.unpackbits(d ^ od).sum() <= 2
    ctx.close()


def test_shard_equivalence(pkg):
    """Frames split over 2 'ranks' with a one-frame overlap give exactly the units of the unsharded run (no collective)."""
    sh = importlib.import_module("svo_b200.sharding")
    w, h, F, B = 640, 200, 9, 4
    left, right = pkg.synth_sequence(77, 0, F, w, h)
    fields = ("kps_left", "n_left", "n_query", "stereo_idx", "temporal_idx", "temporal_dist", "n_seed", "lk_p3", "lk_status",
              "n_orb_matched", "orb_t2_left", "n_lk_matched", "lkf_t2_left")

    def run(f0, n, device=0):
        ctx = pkg.Context(device, w, h, 2 * (B + 1))
        ctx.orb_configure(400, 1.2, 8, 20, 7)
        ctx.frontend_configure(w, h, pkg.MODE_ORB | pkg.MODE_LK, B)
        kc, pc = ctx.orb_capacity(w, h), max(1024, w * h // 64)
        units = {}
        for b0, bn in sh.batches(f0, n, B):
            out = pkg.FrontendBuffers(bn, kc, pc, fields=fields)
            ctx.frontend_batch(np.ascontiguousarray(left[b0:b0 + bn]), np.ascontiguousarray(right[b0:b0 + bn]), out)
            for i in range(bn):
                if b0 + i == f0:
                    continue  # the overlap frame only warms up "last_frame_"
                nq, ns, nm, nk = out.n_query[i], out.n_seed[i], out.n_orb_matched[i], out.n_lk_matched[i]
                units[b0 + i - 1] = (out.kps_left[i, :out.n_left[i]].copy(), out.stereo_idx[i, :nq].copy(), out.temporal_idx[i, :nq].copy(),
                                     out.temporal_dist[i, :nq].copy(), out.lk_p3[i, :ns].copy(), out.lk_status[i, :, :ns].copy(),
                                     out.orb_t2_left[i, :nm].copy(), out.lkf_t2_left[i, :nk].copy())
        ctx.close()
        return units

    whole = run(0, F)
    assert sorted(whole) == list(range(F - 1))
    import torch
    ndev = torch.cuda.device_count()
    # two ranks on ONE device, and -- when the box has more than one GPU -- one rank per device (the real sharded layout; the
    # C5 line of bench.py --config c5 does the same across processes and compares a hash of per-unit signatures)
    layouts = [[0, 0]] + ([[0, 1]] if ndev >= 2 else []) + ([[0, 1, 2, 3]] if ndev >= 4 else [])
    for devices in layouts:
        merged = {}
        for r, dev in enumerate(devices):
            f0, n = sh.shard_units(F, len(devices), r)
            merged.update(run(f0, n, dev))
        assert sorted(merged) == sorted(whole)
        for u in whole:
            for a, b in zip(whole[u], merged[u]):
                assert np.array_equal(a, b), (devices, u)
    print(f"shard equivalence checked on device layouts {layouts} ({ndev} GPUs visible)")


def test_error_paths(pkg):
    L = pkg.lib()
    ctx = pkg.Context(0, 640, 480, 2)
    img = np.zeros((480, 640), np.uint8)
    with pytest.raises(pkg.SvoError) as e:  # not configured
        ctx.orb_extract(img)
    assert e.value.code == pkg.SVO_ERR_NOT_CONFIGURED
    ctx.orb_configure(500, 1.2, 8, 20, 7)
    with pytest.raises(pkg.SvoError) as e:  # image larger than the context
        ctx.orb_extract(np.zeros((481, 640), np.uint8))
    assert e.value.code == pkg.SVO_ERR_INVALID
    kps = np.zeros(10, pkg.KP_DTYPE)
    desc = np.zeros((10, 32), np.uint8)
    n = C.c_int()
    left, _ = pkg.synth_stereo_frame(1, 0, 640, 480)
    rc = L.svo_orb_extract(ctx._h, left.ctypes.data_as(C.c_void_p), 640, 480, C.c_size_t(640), kps.ctypes.data_as(C.c_void_p),
                           desc.ctypes.data_as(C.c_void_p), 10, C.byref(n))
    assert rc == pkg.SVO_ERR_CAPACITY  # never silently truncated
    with pytest.raises(pkg.SvoError):
        ctx.orb_configure(500, 1.0, 8, 20, 7)  # scaleFactor must be > 1
    with pytest.raises(pkg.SvoError):
        ctx.lk_track(left, left, np.zeros((4, 2), np.float32), win=15)
    with pytest.raises(pkg.SvoError):
        ctx.fast_detect(left, 20, nms=False)
    with pytest.raises(pkg.SvoError):
        ctx.frontend_configure(640, 480, pkg.MODE_ORB, 4)  # needs max_images >= 2 * (4 + 1)
    ctx.frontend_configure(320, 240, pkg.MODE_LK, 0 + 1) if False else None
    small = pkg.Context(0, 640, 480, 4)
    small.orb_configure(300, 1.2, 8, 20, 7)
    small.frontend_configure(640, 480, pkg.MODE_ORB, 1)
    with pytest.raises(pkg.SvoError):
        small.frontend_run()  # nothing uploaded
    with pytest.raises(pkg.SvoError):
        small.frontend_download(None)  # nothing to deliver
    small.close()
    ctx.close()


def test_frontend_is_deterministic_at_full_size(pkg):
    """KITTI-sized batch run three times from reset: every output byte identical.  The kernels use shared-memory atomics whose
    ORDER varies between runs (FAST queues, corner lists, per-cell counters); no output may depend on that order."""
    w, h, B = 1241, 376, 12
    left, right = pkg.synth_sequence(1000, 9, B, w, h)  # crosses the canvas re-seed at frame 16
    ctx = pkg.Context(0, w, h, 2 * (B + 1))
    ctx.orb_configure(1000, 1.2, 8, 20, 7)
    ctx.frontend_configure(w, h, pkg.MODE_ORB | pkg.MODE_LK, B)
    kc, pc = ctx.orb_capacity(w, h), max(1024, w * h // 64)
    runs = []
    for _ in range(3):
        ctx.frontend_reset()
        out = pkg.FrontendBuffers(B, kc, pc)
        ctx.frontend_batch(left, right, out)
        runs.append(out)
    # valid prefix of every per-frame array (entries behind the counts are not part of the result)
    count_of = dict(kps_left="n_left", desc_left="n_left", kps_right="n_right", desc_right="n_right", stereo_idx="n_query", stereo_dist="n_query",
                    temporal_idx="n_query", temporal_dist="n_query", fast_kps="n_fast", lk_p1="n_seed", lk_p2="n_seed", lk_p3="n_seed",
                    lk_p0r="n_seed", orb_t2_left="n_orb_matched", orb_t1_left="n_orb_matched", orb_t1_right="n_orb_matched",
                    lkf_t1_left="n_lk_matched", lkf_t1_right="n_lk_matched", lkf_t2_right="n_lk_matched", lkf_t2_left="n_lk_matched")
    a = runs[0]
    for b in runs[1:]:
        for n in ("n_left", "n_right", "n_query", "n_fast", "n_seed", "n_orb_matched", "n_lk_matched"):
            assert np.array_equal(getattr(a, n), getattr(b, n)), n
        for name, cnt in count_of.items():
            for i in range(B):
                k = int(getattr(a, cnt)[i])
                assert np.array_equal(getattr(a, name)[i, :k], getattr(b, name)[i, :k]), (name, i)
        for i in range(B):
            k = int(a.n_seed[i])
            assert np.array_equal(a.lk_status[i, :, :k], b.lk_status[i, :, :k]), ("lk_status", i)
    assert int(runs[0].n_orb_matched[1:].min()) > 50 and int(runs[0].n_lk_matched[1:].min()) > 200
    ctx.close()


def test_long_sequence_is_independent_of_batching(pkg):
    """BASELINE-sized frames, a few hundred of them (several canvas re-seeds): the units of a sequence do not depend on how it is
    cut into batches (128 frames per batch as in bench.py vs 48): per-frame counts and order-sensitive checksums of every list."""
    import zlib
    w, h, F = 1241, 376, 288
    left, right = pkg.synth_sequence(1000, 0, F, w, h)

    def run(B):
        ctx = pkg.Context(0, w, h, 2 * (B + 1))
        ctx.orb_configure(1000, 1.2, 8, 20, 7)
        ctx.frontend_configure(w, h, pkg.MODE_ORB | pkg.MODE_LK, B)
        kc, pc = ctx.orb_capacity(w, h), max(1024, w * h // 64)
        rows = []
        for f0 in range(0, F, B):
            n = min(B, F - f0)
            out = pkg.FrontendBuffers(n, kc, pc, fields=("kps_left", "desc_left", "n_left", "n_query", "temporal_idx", "n_seed", "lk_p3", "lk_status",
                                                         "n_orb_matched", "orb_t2_left", "n_lk_matched", "lkf_t2_left"))
            ctx.frontend_batch(left[f0:f0 + n], right[f0:f0 + n], out)
            for i in range(n):
                nl, nq, ns, nm, nk = (int(x[i]) for x in (out.n_left, out.n_query, out.n_seed, out.n_orb_matched, out.n_lk_matched))
                crc = 0
                for a in (out.kps_left[i, :nl], out.desc_left[i, :nl], out.temporal_idx[i, :nq], out.lk_p3[i, :ns], out.lk_status[i, :, :ns],
                          out.orb_t2_left[i, :nm], out.lkf_t2_left[i, :nk]):
                    crc = zlib.crc32(np.ascontiguousarray(a).tobytes(), crc)
                rows.append((nl, nq, ns, nm, nk, crc))
        ctx.close()
        return rows

    a, b = run(128), run(48)
    assert len(a) == F and a == b
    assert min(r[3] for r in a[1:]) > 20 and min(r[4] for r in a[1:]) > 200


def test_match_random_sizes_against_oracle(pkg, orc):
    """The tcgen05 Hamming match on 120 random (query, train) sizes around its tile edges (128 queries / 256 train descriptors per tile),
    with duplicated train descriptors (ties) and near-duplicates: indices and distances bit-exact, first minimum wins."""
    rng = np.random.default_rng(4242)
    ctx = pkg.Context(0, 64, 64, 2)
    sizes = [(1, 1), (127, 255), (128, 256), (129, 257), (1, 513), (1000, 1), (300, 0)] + [(int(rng.integers(1, 1500)), int(rng.integers(1, 2200))) for _ in range(113)]
    for nq, nt in sizes:
        q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
        t = rng.integers(0, 256, (max(nt, 1), 32), dtype=np.uint8)[:nt]
        if nt >= 4:
            k = min(nq, nt // 2)
            t[nt // 2:nt // 2 + k // 2] = q[:k // 2]            # exact copies of queries ...
            t[:k // 4] = q[:k // 4]                              # ... twice (the lower index must win)
            t[nt - 1] = q[nq - 1] ^ np.uint8(1)                  # distance-1 neighbour in the last column
        gi, gd = ctx.match_hamming(q, t)
        if nt == 0:
            assert (gi == -1).all()
            continue
        oi, od = orc.match_hamming(q, t)
        assert np.array_equal(gi, oi) and np.array_equal(gd, od), (nq, nt)
    ctx.close()
