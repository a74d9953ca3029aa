"""stereo-visual-odometry_b200 -- B200-native stereo visual-odometry front-end.

This package is a thin ctypes view of the C ABI in include/svo_cuda.h (libsvo_b200.so: hand-written sm_100a CUDA
kernels behind an extern "C" layer).  It exists for the tests and bench.py; the product boundary is the C ABI and the
drop-in C++ classes under host/.  There is no CPU fallback: if the shared library is missing this module raises, and
without a CUDA device `Context(...)` raises.

The directory name is not a valid Python identifier; import it with `__graft_entry__.load_package()` (registered as
`svo_b200`).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsvo_b200.so")
SYNTH_LIB_PATH = os.path.join(_HERE, "libsvo_synth.so")  # host-only generator of the synthetic sequences (no CUDA)

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28

SVO_OK, SVO_ERR_INVALID, SVO_ERR_CUDA, SVO_ERR_CAPACITY, SVO_ERR_NOT_CONFIGURED = 0, -1, -2, -3, -4
MODE_ORB, MODE_LK = 1, 2
STAGES = ["pyramid", "fast_cells", "compact", "octree", "describe", "match", "fast_full", "lk_pyramid", "lk_track1", "lk_track2",
          "lk_track3", "lk_track4", "pnp"]
TIMING_SLOTS = 64

# every symbol include/svo_cuda.h declares (tests check the library exports all of them)
ABI_SYMBOLS = [
    "svo_ctx_create", "svo_ctx_destroy", "svo_last_error", "svo_abi_version", "svo_ctx_set_stream", "svo_ctx_synchronize",
    "svo_ctx_launch_count", "svo_ctx_enable_stage_timing", "svo_ctx_set_timing_slot", "svo_ctx_stage_ms", "svo_orb_configure", "svo_orb_tables",
    "svo_orb_capacity", "svo_orb_level_size", "svo_orb_extract", "svo_orb_extract_batch", "svo_orb_read_level",
    "svo_orb_read_candidates", "svo_fast_detect", "svo_match_hamming", "svo_lk_track", "svo_frontend_configure",
    "svo_frontend_reset", "svo_frontend_set_match_error", "svo_frontend_set_overlap", "svo_frontend_set_lk_err", "svo_frontend_set_projection",
    "svo_triangulate_points", "svo_frontend_upload", "svo_frontend_wait_upload", "svo_frontend_upload_dev", "svo_frontend_run", "svo_frontend_download", "svo_frontend_check",
    "svo_frontend_discard", "svo_frontend_batch", "svo_source_open", "svo_source_frame_count", "svo_source_is_pinned", "svo_source_next",
    "svo_source_error", "svo_source_close", "svo_write_pgm", "svo_write_png", "svo_orb_match_filter", "svo_lk_chain_filter", "svo_guard_check", "svo_guard_selftest",
    "svo_pnp_ransac", "svo_frontend_set_pnp",
]
SYNTH_SYMBOLS = ["svo_synth_stereo_frame"]  # exported by libsvo_synth.so


class SvoError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"svo error {code}: {msg}")
        self.code = code


class FrontendOut(C.Structure):
    _fields_ = [("kp_cap", C.c_int32), ("pt_cap", C.c_int32)] + [(n, C.c_void_p) for n in (
        "kps_left", "desc_left", "n_left", "kps_right", "desc_right", "n_right", "n_query", "stereo_idx", "stereo_dist",
        "temporal_idx", "temporal_dist", "fast_kps", "n_fast", "n_seed", "lk_p1", "lk_p2", "lk_p3", "lk_p0r", "lk_status",
        "orb_t2_left", "orb_t1_left", "orb_t1_right", "n_orb_matched", "lkf_t1_left", "lkf_t1_right", "lkf_t2_right", "lkf_t2_left",
        "n_lk_matched", "orb_points3d", "lk_points3d", "lk_err", "orb_pose", "orb_pnp_info", "orb_inlier_mask", "lk_pose", "lk_pnp_info",
        "lk_inlier_mask")]


_lib = None
_synth = None


def guard_check():
    """Debug (SVO_GUARD=1 in the environment before the library loads): guard-band bytes overwritten so far; -1 if guards are off."""
    return int(lib().svo_guard_check())


def guard_selftest():
    """Debug: a deliberate 3-byte overrun of a scratch buffer as seen by the guard bands (3 with SVO_GUARD=1, -1 without)."""
    return int(lib().svo_guard_selftest())


def synth_lib():
    """Load libsvo_synth.so: the host-only synthetic-sequence generator (the CPU reference arm uses it without the CUDA library)."""
    global _synth
    if _synth is None:
        if not os.path.exists(SYNTH_LIB_PATH):
            raise ImportError(f"{SYNTH_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(SYNTH_LIB_PATH)
        L.svo_synth_stereo_frame.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
        _synth = L
    return _synth


def lib():
    """Load libsvo_b200.so (fails loudly if the CUDA extension has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.svo_last_error.restype = C.c_char_p
        L.svo_last_error.argtypes = [C.c_void_p]
        L.svo_ctx_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int]
        L.svo_ctx_destroy.argtypes = [C.c_void_p]
        L.svo_ctx_destroy.restype = None
        L.svo_guard_check.restype = C.c_longlong
        L.svo_guard_check.argtypes = []
        L.svo_guard_selftest.restype = C.c_longlong
        L.svo_guard_selftest.argtypes = []
        L.svo_ctx_launch_count.restype = C.c_int64
        L.svo_ctx_launch_count.argtypes = [C.c_void_p]
        L.svo_ctx_stage_ms.restype = C.c_float
        L.svo_ctx_stage_ms.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.svo_orb_configure.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.svo_source_open.argtypes = [C.POINTER(C.c_void_p), C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.svo_source_frame_count.argtypes = [C.c_void_p]
        L.svo_source_is_pinned.argtypes = [C.c_void_p]
        L.svo_source_next.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
        L.svo_source_error.restype = C.c_char_p
        L.svo_source_error.argtypes = [C.c_void_p]
        L.svo_source_close.argtypes = [C.c_void_p]
        L.svo_source_close.restype = None
        L.svo_write_pgm.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t]
        L.svo_write_png.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int]
        L.svo_lk_track.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]
        L.svo_pnp_ransac.argtypes = [C.c_void_p] + [C.c_void_p] * 4 + [C.c_int, C.c_int, C.c_int, C.c_float, C.c_double] + [C.c_void_p] * 7
        L.svo_frontend_set_pnp.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_double]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def synth_stereo_frame(seed, frame, w, h):
    """Frame `frame` of the deterministic synthetic stereo sequence `seed`. Host code, no GPU needed."""
    left = np.empty((h, w), np.uint8)
    right = np.empty((h, w), np.uint8)
    rc = synth_lib().svo_synth_stereo_frame(seed, frame, w, h, _p(left), _p(right), w)
    if rc != 0:
        raise SvoError(rc, "svo_synth_stereo_frame")
    return left, right


def synth_sequence(seed, first, count, w, h):
    left = np.empty((count, h, w), np.uint8)
    right = np.empty((count, h, w), np.uint8)
    for i in range(count):
        rc = synth_lib().svo_synth_stereo_frame(seed, first + i, w, h, _p(left[i]), _p(right[i]), w)
        if rc != 0:
            raise SvoError(rc, "svo_synth_stereo_frame")
    return left, right


def write_sequence(directory, left, right, first_index=0, fmt="png", zlib_level=1):
    """left / right (n, h, w) uint8 -> <directory>/image_0|1/%06d.<fmt> (the reference's dataset layout, System.cpp:77)."""
    L = lib()
    n, h, w = left.shape
    for side, imgs in ((0, left), (1, right)):
        os.makedirs(os.path.join(directory, f"image_{side}"), exist_ok=True)
        for i in range(n):
            path = os.path.join(directory, f"image_{side}", f"{first_index + i:06d}.{fmt}").encode()
            img = np.ascontiguousarray(imgs[i])
            rc = L.svo_write_png(path, _p(img), w, h, w, zlib_level) if fmt == "png" else L.svo_write_pgm(path, _p(img), w, h, w)
            if rc != 0:
                raise SvoError(rc, f"cannot write {path.decode()}")


class FrameSource:
    """svo_source: decoder threads filling a pinned ring of batches from a KITTI-layout directory (System::NextFrame_kitti)."""

    def __init__(self, directory, w, h, batch, ring_batches=3, threads=8, first_index=0):
        self._L = lib()
        hnd = C.c_void_p()
        rc = self._L.svo_source_open(C.byref(hnd), directory.encode(), first_index, w, h, batch, ring_batches, threads)
        if rc != 0:
            raise SvoError(rc, "svo_source_open failed")
        self._h, self.w, self.h, self.batch = hnd, w, h, batch

    def frame_count(self):
        return int(self._L.svo_source_frame_count(self._h))

    def is_pinned(self):
        return bool(self._L.svo_source_is_pinned(self._h))

    def next_ptr(self):
        """-> (left address, right address, n_frames); n_frames == 0 at the end of the run"""
        l, r, n = C.c_void_p(), C.c_void_p(), C.c_int()
        rc = self._L.svo_source_next(self._h, C.byref(l), C.byref(r), C.byref(n))
        if rc != 0:
            raise SvoError(rc, self._L.svo_source_error(self._h).decode())
        return l.value, r.value, n.value

    def next(self):
        """-> (left, right) numpy views (n, h, w) into the ring, valid until the next call; None at the end of the run"""
        l, r, n = self.next_ptr()
        if n == 0:
            return None
        shape = (n, self.h, self.w)
        mk = lambda a: np.ctypeslib.as_array(C.cast(a, C.POINTER(C.c_uint8)), shape=shape)
        return mk(l), mk(r)

    def close(self):
        if getattr(self, "_h", None):
            self._L.svo_source_close(self._h)
            self._h = None

    def __del__(self):
        self.close()


class Context:
    """One svo_ctx: one device, one stream, not re-entrant (mirrors the reference's per-object scratch)."""

    def __init__(self, device=0, max_w=1241, max_h=376, max_images=2):
        self._L = lib()
        h = C.c_void_p()
        rc = self._L.svo_ctx_create(C.byref(h), device, max_w, max_h, max_images)
        if rc != 0:
            raise SvoError(rc, "svo_ctx_create failed (no CUDA device? there is no CPU fallback)")
        self._h = h
        self.max_images = max_images
        self.nlevels = 0

    def close(self):
        if getattr(self, "_h", None):
            self._L.svo_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise SvoError(rc, self._L.svo_last_error(self._h).decode())

    # ---- context
    def set_stream(self, cuda_stream_handle):
        self._check(self._L.svo_ctx_set_stream(self._h, C.c_void_p(cuda_stream_handle)))

    def synchronize(self):
        self._check(self._L.svo_ctx_synchronize(self._h))

    def launch_count(self):
        return int(self._L.svo_ctx_launch_count(self._h))

    def enable_stage_timing(self, on=True):
        self._check(self._L.svo_ctx_enable_stage_timing(self._h, int(on)))

    def set_timing_slot(self, slot):
        self._check(self._L.svo_ctx_set_timing_slot(self._h, slot))

    def stage_ms(self, slot=0):
        return {name: float(self._L.svo_ctx_stage_ms(self._h, slot, i)) for i, name in enumerate(STAGES)}

    # ---- ORB extractor (ORBextractor.h:33-75)
    def orb_configure(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self._check(self._L.svo_orb_configure(self._h, nfeatures, scale_factor, nlevels, ini_th, min_th))
        self.nlevels = nlevels

    def orb_tables(self):
        n = self.nlevels
        arrs = [np.empty(n, np.float32) for _ in range(4)] + [np.empty(n, np.int32)]
        self._check(self._L.svo_orb_tables(self._h, *[_p(a) for a in arrs]))
        return dict(zip(("scale", "inv_scale", "sigma2", "inv_sigma2", "features_per_level"), arrs))

    def orb_capacity(self, w, h):
        cap = self._L.svo_orb_capacity(self._h, w, h)
        if cap < 0:
            self._check(cap)
        return cap

    def orb_level_size(self, w, h, level):
        lw, lh = C.c_int(), C.c_int()
        self._check(self._L.svo_orb_level_size(self._h, w, h, level, C.byref(lw), C.byref(lh)))
        return lw.value, lh.value

    def orb_extract(self, img):
        """ORBextractor::operator() on one image -> (keypoints[KP_DTYPE], descriptors[N,32])."""
        k, d = self.orb_extract_batch([img])
        return k[0], d[0]

    def orb_extract_batch(self, imgs):
        imgs = [np.ascontiguousarray(i, np.uint8) for i in imgs]
        n = len(imgs)
        if n == 0:
            return [], []
        h, w = imgs[0].shape
        if w == 0 or h == 0:
            return [np.zeros(0, KP_DTYPE) for _ in imgs], [np.zeros((0, 32), np.uint8) for _ in imgs]
        assert all(i.shape == (h, w) for i in imgs)
        cap = self.orb_capacity(w, h)
        kps = np.zeros((n, cap), KP_DTYPE)
        desc = np.zeros((n, cap, 32), np.uint8)
        nout = np.zeros(n, np.int32)
        ptrs = (C.c_void_p * n)(*[i.ctypes.data for i in imgs])
        self._check(self._L.svo_orb_extract_batch(self._h, ptrs, n, w, h, C.c_size_t(imgs[0].strides[0]), _p(kps), _p(desc), cap, _p(nout)))
        return [kps[i, :nout[i]].copy() for i in range(n)], [desc[i, :nout[i]].copy() for i in range(n)]

    def orb_read_level(self, slot, level):
        w, h = C.c_int(), C.c_int()
        self._check(self._L.svo_orb_read_level(self._h, slot, level, None, C.c_size_t(0), C.byref(w), C.byref(h)))
        out = np.empty((h.value, w.value), np.uint8)
        self._check(self._L.svo_orb_read_level(self._h, slot, level, _p(out), C.c_size_t(w.value), C.byref(w), C.byref(h)))
        return out

    def orb_read_candidates(self, slot, level):
        n = C.c_int()
        self._check(self._L.svo_orb_read_candidates(self._h, slot, level, None, 0, C.byref(n)))
        out = np.zeros(max(n.value, 1), KP_DTYPE)
        self._check(self._L.svo_orb_read_candidates(self._h, slot, level, _p(out), max(n.value, 1), C.byref(n)))
        return out[:n.value]

    # ---- full-frame FAST (tracking.cpp:101)
    def fast_detect(self, img, threshold=20, nms=True, cap=None):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        cap = cap or max(1024, w * h // 16)
        out = np.zeros(cap, KP_DTYPE)
        n = C.c_int()
        self._check(self._L.svo_fast_detect(self._h, _p(img), w, h, C.c_size_t(img.strides[0]), threshold, int(nms), _p(out), cap, C.byref(n)))
        return out[:n.value].copy()

    # ---- BF Hamming (tracking.cpp:543-544)
    def match_hamming(self, q, t):
        q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
        t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
        idx = np.full(len(q), -1, np.int32)
        dist = np.full(len(q), -1, np.int32)
        self._check(self._L.svo_match_hamming(self._h, _p(q), len(q), _p(t), len(t), _p(idx), _p(dist)))
        return idx, dist

    # ---- pyramidal LK (tracking.cpp:593-618)
    def lk_track(self, prev, nxt, pts, win=21, max_level=3, max_iter=30, eps=0.01, min_eig=1e-3):
        prev = np.ascontiguousarray(prev, np.uint8)
        nxt = np.ascontiguousarray(nxt, np.uint8)
        assert prev.shape == nxt.shape
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
        n = len(pts)
        out = np.zeros((n, 2), np.float32)
        status = np.zeros(n, np.uint8)
        err = np.zeros(n, np.float32)
        h, w = prev.shape
        self._check(self._L.svo_lk_track(self._h, _p(prev), _p(nxt), w, h, prev.strides[0], _p(pts), n, _p(out), _p(status), _p(err),
                                         win, max_level, max_iter, eps, min_eig))
        return out, status, err

    # ---- the Tracking post-filters on the device (tracking.cpp:546-577, 623-660)
    def orb_match_filter(self, last_left, last_right, cur_left, idx1, dist1, idx2, dist2, fme=3.0):
        """-> (t2_left, t1_left, t1_right) point lists [n, 2]"""
        ll, lr, cl = (np.ascontiguousarray(a, KP_DTYPE) for a in (last_left, last_right, cur_left))
        i1, d1, i2, d2 = (np.ascontiguousarray(a, np.int32) for a in (idx1, dist1, idx2, dist2))
        cap = max(1, min(len(ll), len(lr), len(cl)))
        o = [np.zeros((cap, 2), np.float32) for _ in range(3)]
        n = C.c_int()
        self._check(self._L.svo_orb_match_filter(self._h, _p(ll), len(ll), _p(lr), len(lr), _p(cl), len(cl), _p(i1), _p(d1), _p(i2), _p(d2),
                                                 C.c_double(fme), _p(o[0]), _p(o[1]), _p(o[2]), cap, C.byref(n)))
        return tuple(a[:n.value].copy() for a in o)

    def lk_chain_filter(self, p0, p1, p2, p3, p0r, s0, s1, s2, s3, fme=3.0):
        """-> the five compacted point lists"""
        pts = [np.ascontiguousarray(a, np.float32).reshape(-1, 2).copy() for a in (p0, p1, p2, p3, p0r)]
        st = [np.ascontiguousarray(a, np.uint8).ravel() for a in (s0, s1, s2, s3)]
        n = C.c_int()
        self._check(self._L.svo_lk_chain_filter(self._h, len(pts[0]), *[_p(a) for a in pts], *[_p(a) for a in st], C.c_double(fme), C.byref(n)))
        return tuple(a[:n.value] for a in pts)

    # ---- throughput mode
    def frontend_configure(self, w, h, mode, max_frames):
        self._check(self._L.svo_frontend_configure(self._h, w, h, mode, max_frames))
        self.fe = dict(w=w, h=h, mode=mode, max_frames=max_frames, kp_cap=self.orb_capacity(w, h) if mode & MODE_ORB else 0)

    def frontend_set_match_error(self, fme):
        self._check(self._L.svo_frontend_set_match_error(self._h, C.c_double(fme)))

    def frontend_set_overlap(self, enable):
        self._check(self._L.svo_frontend_set_overlap(self._h, int(bool(enable))))

    def frontend_set_lk_err(self, enable):
        self._check(self._L.svo_frontend_set_lk_err(self._h, int(bool(enable))))

    def frontend_check(self):
        """Raises SvoError(SVO_ERR_CAPACITY) if a run retired with frontend_download(None) overflowed a device capacity."""
        self._check(self._L.svo_frontend_check(self._h))

    def frontend_discard(self):
        self._check(self._L.svo_frontend_discard(self._h))

    def frontend_set_projection(self, p1, p2):
        p1 = np.ascontiguousarray(p1, np.float64).reshape(12)
        p2 = np.ascontiguousarray(p2, np.float64).reshape(12)
        self._check(self._L.svo_frontend_set_projection(self._h, _p(p1), _p(p2)))

    def frontend_set_pnp(self, enable=True, iterations=500, reproj_error=0.5, confidence=float(np.float32(0.99))):
        """Tracking::iterationsCount_ / reprojectionError_ / confidence_ (config/default.yaml:80-82; confidence_ is a float there)."""
        self._check(self._L.svo_frontend_set_pnp(self._h, int(bool(enable)), int(iterations), C.c_float(reproj_error), C.c_double(confidence)))

    def pnp_ransac(self, K, pts3d, pts2d, counts=None, iterations=500, reproj_error=0.5, confidence=0.99):
        """cv::solvePnPRansac (tracking.cpp:486-488) for one list ([n,3] / [n,2]) or a batch ([b,cap,3] / [b,cap,2] + counts[b]).
        Returns dict(ok, rvec, tvec, R, mask, n_inliers, best_iter, iters); arrays carry a leading batch axis for batch input."""
        K = np.ascontiguousarray(K, np.float64).reshape(9)
        p3 = np.ascontiguousarray(pts3d, np.float32)
        p2 = np.ascontiguousarray(pts2d, np.float32)
        single = p3.ndim == 2
        if single:
            p3, p2 = p3[None], p2[None]
        b, cap = p3.shape[0], p3.shape[1]
        cnt = np.full(b, cap, np.int32) if counts is None else np.ascontiguousarray(counts, np.int32)
        rvec, tvec, R = np.zeros((b, 3)), np.zeros((b, 3)), np.zeros((b, 9))  # noqa
        mask = np.zeros((b, max(cap, 1)), np.uint8)
        ninl, info, ok = np.zeros(b, np.int32), np.zeros((b, 2), np.int32), np.zeros(b, np.int32)
        self._check(self._L.svo_pnp_ransac(self._h, _p(K), _p(p3), _p(p2), _p(cnt), b, cap, int(iterations), C.c_float(reproj_error),
                                           C.c_double(confidence), _p(rvec), _p(tvec), _p(R), _p(mask), _p(ninl), _p(info), _p(ok)))
        res = dict(ok=ok.astype(bool), rvec=rvec, tvec=tvec, R=R.reshape(b, 3, 3), mask=mask[:, :cap].astype(bool), n_inliers=ninl,
                   best_iter=info[:, 0].copy(), iters=info[:, 1].copy())
        return {k: v[0] for k, v in res.items()} if single else res

    def triangulate_points(self, p1, p2, pts1, pts2):
        """cv::triangulatePoints + convertPointsFromHomogeneous -> (points4d [n,4], points3d [n,3])"""
        p1 = np.ascontiguousarray(p1, np.float64).reshape(12)
        p2 = np.ascontiguousarray(p2, np.float64).reshape(12)
        pts1 = np.ascontiguousarray(pts1, np.float32).reshape(-1, 2)
        pts2 = np.ascontiguousarray(pts2, np.float32).reshape(-1, 2)
        n = len(pts1)
        o4 = np.zeros((n, 4), np.float32)
        o3 = np.zeros((n, 3), np.float32)
        self._check(self._L.svo_triangulate_points(self._h, _p(p1), _p(p2), _p(pts1), _p(pts2), n, _p(o4), _p(o3)))
        return o4, o3

    def frontend_reset(self):
        self._check(self._L.svo_frontend_reset(self._h))

    def frontend_upload(self, left, right):
        assert left.shape == right.shape and left.dtype == np.uint8 and left.flags.c_contiguous and right.flags.c_contiguous
        self._check(self._L.svo_frontend_upload(self._h, _p(left), _p(right), left.shape[0], C.c_size_t(left.strides[1])))
        self._n_frames = left.shape[0]

    def frontend_upload_ptr(self, left_ptr, right_ptr, n_frames, stride):
        self._check(self._L.svo_frontend_upload(self._h, C.c_void_p(left_ptr), C.c_void_p(right_ptr), n_frames, C.c_size_t(stride)))
        self._n_frames = n_frames

    def frontend_wait_upload(self):
        self._check(self._L.svo_frontend_wait_upload(self._h))

    def frontend_upload_dev(self, left_ptr, right_ptr, n_frames, stride):
        """left_ptr / right_ptr: device addresses of n_frames contiguous images (rows of `stride` bytes)."""
        self._check(self._L.svo_frontend_upload_dev(self._h, C.c_void_p(left_ptr), C.c_void_p(right_ptr), n_frames, C.c_size_t(stride)))
        self._n_frames = n_frames

    def frontend_run(self):
        self._check(self._L.svo_frontend_run(self._h))

    def frontend_download(self, out):
        self._check(self._L.svo_frontend_download(self._h, C.byref(out.struct) if out is not None else None))

    def frontend_batch(self, left, right, out):
        assert left.shape == right.shape and left.dtype == np.uint8 and left.flags.c_contiguous and right.flags.c_contiguous
        self._check(self._L.svo_frontend_batch(self._h, _p(left), _p(right), left.shape[0], C.c_size_t(left.strides[1]), C.byref(out.struct)))


class FrontendBuffers:
    """Host-side output arrays of svo_frontend_download / svo_frontend_batch (numpy-owned; `fields` selects what is copied)."""

    ALL = ("kps_left", "desc_left", "n_left", "kps_right", "desc_right", "n_right", "n_query", "stereo_idx", "stereo_dist",
           "temporal_idx", "temporal_dist", "fast_kps", "n_fast", "n_seed", "lk_p1", "lk_p2", "lk_p3", "lk_p0r", "lk_status",
           "orb_t2_left", "orb_t1_left", "orb_t1_right", "n_orb_matched", "lkf_t1_left", "lkf_t1_right", "lkf_t2_right", "lkf_t2_left",
           "n_lk_matched", "orb_points3d", "lk_points3d", "lk_err", "orb_pose", "orb_pnp_info", "orb_inlier_mask", "lk_pose", "lk_pnp_info",
           "lk_inlier_mask")

    def __init__(self, n_frames, kp_cap, pt_cap, fields=None, alloc=None):
        fields = self.ALL if fields is None else fields
        self.n_frames, self.kp_cap, self.pt_cap = n_frames, kp_cap, pt_cap
        F, K, P = n_frames, max(kp_cap, 1), max(pt_cap, 1)
        shapes = dict(kps_left=((F, K), KP_DTYPE), desc_left=((F, K, 32), np.uint8), n_left=((F,), np.int32),
                      kps_right=((F, K), KP_DTYPE), desc_right=((F, K, 32), np.uint8), n_right=((F,), np.int32),
                      n_query=((F,), np.int32), stereo_idx=((F, K), np.int32), stereo_dist=((F, K), np.int32),
                      temporal_idx=((F, K), np.int32), temporal_dist=((F, K), np.int32),
                      fast_kps=((F, P), KP_DTYPE), n_fast=((F,), np.int32), n_seed=((F,), np.int32),
                      lk_p1=((F, P, 2), np.float32), lk_p2=((F, P, 2), np.float32), lk_p3=((F, P, 2), np.float32),
                      lk_p0r=((F, P, 2), np.float32), lk_status=((F, 4, P), np.uint8),
                      orb_t2_left=((F, K, 2), np.float32), orb_t1_left=((F, K, 2), np.float32), orb_t1_right=((F, K, 2), np.float32),
                      n_orb_matched=((F,), np.int32), lkf_t1_left=((F, P, 2), np.float32), lkf_t1_right=((F, P, 2), np.float32),
                      lkf_t2_right=((F, P, 2), np.float32), lkf_t2_left=((F, P, 2), np.float32), n_lk_matched=((F,), np.int32),
                      orb_points3d=((F, K, 3), np.float32), lk_points3d=((F, P, 3), np.float32), lk_err=((F, 4, P), np.float32),
                      orb_pose=((F, 16), np.float64), orb_pnp_info=((F, 4), np.int32), orb_inlier_mask=((F, K), np.uint8),
                      lk_pose=((F, 16), np.float64), lk_pnp_info=((F, 4), np.int32), lk_inlier_mask=((F, P), np.uint8))
        self.pnp_enabled = False  # set by the caller when svo_frontend_set_pnp() is on (copied_bytes counts the pose outputs then)
        self.struct = FrontendOut()
        self.struct.kp_cap = kp_cap
        self.struct.pt_cap = pt_cap
        self.bytes = 0
        for name in self.ALL:
            if name in fields:
                shape, dt = shapes[name]
                arr = alloc(shape, dt) if alloc else np.zeros(shape, dt)
                arr[...] = 0
                setattr(self, name, arr)
                setattr(self.struct, name, arr.ctypes.data)
                self.bytes += arr.nbytes
            else:
                setattr(self, name, None)
                setattr(self.struct, name, None)

    def copied_bytes(self):
        """Bytes svo_frontend_download moved device->host for the last call (per-frame widths are the batch maxima)."""
        F, total = self.n_frames, 0

        def mx(a):
            return int(a.max()) if a is not None and len(a) else 0
        ml, mr, mq, mf, ms = mx(self.n_left), mx(self.n_right), mx(self.n_query), mx(self.n_fast), mx(self.n_seed)
        mo, mk = mx(self.n_orb_matched), mx(self.n_lk_matched)
        for name, width in (("kps_left", ml * 28), ("desc_left", ml * 32), ("kps_right", mr * 28), ("desc_right", mr * 32),
                            ("stereo_idx", mq * 4), ("stereo_dist", mq * 4), ("temporal_idx", mq * 4), ("temporal_dist", mq * 4),
                            ("fast_kps", mf * 28), ("lk_p1", ms * 8), ("lk_p2", ms * 8), ("lk_p3", ms * 8), ("lk_p0r", ms * 8),
                            ("lk_status", ms * 4), ("orb_t2_left", mo * 8), ("orb_t1_left", mo * 8), ("orb_t1_right", mo * 8),
                            ("lkf_t1_left", mk * 8), ("lkf_t1_right", mk * 8), ("lkf_t2_right", mk * 8), ("lkf_t2_left", mk * 8),
                            ("orb_points3d", mo * 12), ("lk_points3d", mk * 12), ("lk_err", ms * 16)):
            if getattr(self, name) is not None:
                total += width * F
        if self.pnp_enabled:
            for name, width in (("orb_pose", 128), ("orb_pnp_info", 16), ("orb_inlier_mask", mo), ("lk_pose", 128), ("lk_pnp_info", 16),
                                ("lk_inlier_mask", mk)):
                if getattr(self, name) is not None:
                    total += width * F
        return total + 3 * 4 * 2 * (F + 1)  # the count / flag arrays
