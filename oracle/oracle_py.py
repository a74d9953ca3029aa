"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle_prims.cpp", "oracle_orb.cpp", "oracle_capi.cpp", "oracle_pnp.cpp", "oracle.h", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_fast_atan2.restype = C.c_float
        L.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
        L.orc_brief_pattern.restype = C.POINTER(C.c_int8)
        L.orc_orb_create.restype = C.c_void_p
        L.orc_orb_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.orc_orb_destroy.argtypes = [C.c_void_p]
        L.orc_frontend_sequence.restype = C.c_uint64
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _img(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    assert a.ndim == 2
    return a


def resize_linear(src, dw, dh):
    src = _img(src)
    dst = np.empty((dh, dw), np.uint8)
    lib().orc_resize_linear_u8(_p(src), C.c_int(src.shape[1]), C.c_int(src.shape[0]), C.c_size_t(src.strides[0]), _p(dst),
                               C.c_int(dw), C.c_int(dh), C.c_size_t(dw))
    return dst


def fast9(img, threshold, nms=True, cap=1 << 20):
    img = _img(img)
    out = np.zeros(cap, KP_DTYPE)
    n = lib().orc_fast9_detect(_p(img), C.c_int(img.shape[1]), C.c_int(img.shape[0]), C.c_size_t(img.strides[0]),
                               C.c_int(threshold), C.c_int(int(nms)), _p(out), C.c_int(cap))
    assert n <= cap
    return out[:n]


def fast9_score_map(img, min_score):
    img = _img(img)
    out = np.empty(img.shape, np.uint8)
    lib().orc_fast9_score_map(_p(img), C.c_int(img.shape[1]), C.c_int(img.shape[0]), C.c_size_t(img.strides[0]),
                              C.c_int(min_score), _p(out))
    return out


def gaussian_blur7(img):
    img = _img(img)
    out = np.empty_like(img)
    lib().orc_gaussian_blur7(_p(img), C.c_int(img.shape[1]), C.c_int(img.shape[0]), C.c_size_t(img.strides[0]), _p(out),
                             C.c_size_t(img.shape[1]))
    return out


def fast_atan2(y, x):
    y = np.ascontiguousarray(y, np.float32)
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(y)
    lib().orc_fast_atan2_array(_p(y), _p(x), _p(out), C.c_int(y.size))
    return out


def match_hamming(q, t):
    q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
    t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
    idx = np.empty(len(q), np.int32)
    dist = np.empty(len(q), np.int32)
    lib().orc_match_hamming(_p(q), C.c_int(len(q)), _p(t), C.c_int(len(t)), _p(idx), _p(dist))
    return idx, dist


def pyr_down(img):
    img = _img(img)
    h, w = img.shape
    out = np.empty(((h + 1) // 2, (w + 1) // 2), np.uint8)
    lib().orc_pyr_down(_p(img), C.c_int(w), C.c_int(h), C.c_size_t(img.strides[0]), _p(out), C.c_size_t(out.shape[1]))
    return out


def lk_track(prev, nxt, pts, win=21, max_level=3, max_iter=30, eps=0.01, min_eig=1e-3):
    prev, nxt = _img(prev), _img(nxt)
    assert prev.shape == nxt.shape
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
    n = len(pts)
    out = np.zeros((n, 2), np.float32)
    status = np.zeros(n, np.uint8)
    err = np.zeros(n, np.float32)
    lib().orc_lk_track(_p(prev), _p(nxt), C.c_int(prev.shape[1]), C.c_int(prev.shape[0]), C.c_size_t(prev.strides[0]), _p(pts),
                       C.c_int(n), _p(out), _p(status), _p(err), C.c_int(win), C.c_int(max_level), C.c_int(max_iter),
                       C.c_double(eps), C.c_double(min_eig))
    return out, status, err


def orb_tables(nfeatures, scale_factor, nlevels, ini_th=20, min_th=7):
    scale = np.empty(nlevels, np.float32)
    inv = np.empty(nlevels, np.float32)
    fpl = np.empty(nlevels, np.int32)
    umax = np.empty(16, np.int32)
    lib().orc_orb_tables(C.c_int(nfeatures), C.c_float(scale_factor), C.c_int(nlevels), C.c_int(ini_th), C.c_int(min_th),
                         _p(scale), _p(inv), _p(fpl), _p(umax))
    return scale, inv, fpl, umax


def brief_pattern():
    ptr = lib().orc_brief_pattern()
    return np.ctypeslib.as_array(ptr, shape=(1024,)).copy()


def distribute_octree(keys, min_x, max_x, min_y, max_y, n_target):
    keys = np.ascontiguousarray(keys, KP_DTYPE)
    cap = max(16, n_target + 64, 4 * 16)
    out = np.zeros(cap, KP_DTYPE)
    stats = np.zeros(2, np.int32)
    n = lib().orc_distribute_octree(_p(keys), C.c_int(len(keys)), C.c_int(min_x), C.c_int(max_x), C.c_int(min_y),
                                    C.c_int(max_y), C.c_int(n_target), _p(out), C.c_int(cap), _p(stats))
    assert n <= cap
    return out[:n], {"rounds": int(stats[0]), "break_in_tie_group": int(stats[1])}


class OrbOracle:
    """ORBextractor restatement with every intermediate kept (ORBextractor.cpp:990-1055)."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.nlevels = nlevels
        self.nfeatures = nfeatures
        self._h = lib().orc_orb_create(nfeatures, scale_factor, nlevels, ini_th, min_th)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_orb_destroy(self._h)
            self._h = None

    def extract(self, img):
        img = _img(img)
        cap = self.nfeatures + 64 * self.nlevels + 64
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = lib().orc_orb_extract(C.c_void_p(self._h), _p(img), C.c_int(img.shape[1]), C.c_int(img.shape[0]),
                                  C.c_size_t(img.strides[0]), _p(kps), _p(desc), C.c_int(cap))
        assert n <= cap
        return kps[:n].copy(), desc[:n].copy()

    def level(self, l):
        w, h = C.c_int(), C.c_int()
        assert lib().orc_orb_level_size(C.c_void_p(self._h), C.c_int(l), C.byref(w), C.byref(h)) == 0
        out = np.empty((h.value, w.value), np.uint8)
        lib().orc_orb_level_pixels(C.c_void_p(self._h), C.c_int(l), _p(out))
        return out

    def _kps(self, fn, l):
        n = fn(C.c_void_p(self._h), C.c_int(l), None, C.c_int(0))
        out = np.zeros(max(n, 1), KP_DTYPE)
        fn(C.c_void_p(self._h), C.c_int(l), _p(out), C.c_int(max(n, 1)))
        return out[:n]

    def candidates(self, l):
        return self._kps(lib().orc_orb_level_candidates, l)

    def level_keypoints(self, l):
        return self._kps(lib().orc_orb_level_keypoints, l)

    def octree_stats(self, l):
        r, t = C.c_int(), C.c_int()
        lib().orc_orb_octree_stats(C.c_void_p(self._h), C.c_int(l), C.byref(r), C.byref(t))
        return {"rounds": r.value, "break_in_tie_group": t.value}


# ---- the reference's own ORBextractor.cpp, compiled where it lies against oracle/ref_shim (oracle/Makefile, target `ref`) ----
_REF = None
REF_SRC = "/root/reference/src/ORBextractor.cpp"


def build_ref(force=False):
    """oracle/_ref/liborb_ref.so, or None where neither the prebuilt library nor /root/reference exists (e.g. a GPU box
    that received no prebuilt copy).  Never reads /root/reference at run time once the library is built."""
    so = os.path.join(_HERE, "_ref", "liborb_ref.so")
    if os.path.exists(REF_SRC):
        deps = [REF_SRC, "/root/reference/src/tracking.cpp"] + [os.path.join(_HERE, f) for f in (
            "ref_capi.cpp", "ref_tracking_capi.cpp", "ref_shim/shim_cv.hpp", "oracle_prims.cpp", "Makefile")]
        if force or not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
            subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
    return so if os.path.exists(so) else None


def ref_lib():
    global _REF
    if _REF is None:
        so = build_ref()
        if so is None:
            return None
        L = C.CDLL(so)
        L.ref_orb_create.restype = C.c_void_p
        L.ref_orb_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.ref_orb_destroy.argtypes = [C.c_void_p]
        L.ref_orb_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_orb_level.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.ref_set_allocator.restype = C.c_size_t
        L.ref_set_allocator.argtypes = [C.c_int]
        _REF = L
    return _REF


class RefOrb:
    """lzb_vio::ORBextractor as compiled from the reference's source.  monotone=True serves the library's allocations from
    a bump arena (addresses grow with creation order): the allocator under which the octree's pointer tie-break
    (ORBextractor.cpp:636) is reproducible; monotone=False uses the system malloc (heap-history dependent)."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, monotone=True):
        self.L = ref_lib()
        assert self.L is not None, "oracle/_ref/liborb_ref.so is not available"
        self.monotone, self.nfeatures, self.nlevels = monotone, nfeatures, nlevels
        self.L.ref_set_allocator(1 if monotone else 0)
        self._h = self.L.ref_orb_create(nfeatures, scale_factor, nlevels, ini_th, min_th)
        self.L.ref_set_allocator(0)

    def __del__(self):
        if getattr(self, "_h", None):
            self.L.ref_orb_destroy(self._h)
            self._h = None

    def extract(self, img):
        img = _img(img)
        cap = self.nfeatures + 64 * self.nlevels + 64
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        self.L.ref_set_allocator(1 if self.monotone else 0)
        n = self.L.ref_orb_extract(C.c_void_p(self._h), _p(img), img.shape[1], img.shape[0], img.strides[0], _p(kps), _p(desc), cap)
        self.L.ref_set_allocator(0)
        assert 0 <= n <= cap, n
        return kps[:n].copy(), desc[:n].copy()

    def level(self, l):
        w, h = C.c_int(), C.c_int()
        assert self.L.ref_orb_level(C.c_void_p(self._h), l, None, 0, C.byref(w), C.byref(h)) == 0
        out = np.empty((h.value, w.value), np.uint8)
        self.L.ref_orb_level(C.c_void_p(self._h), l, _p(out), out.strides[0], C.byref(w), C.byref(h))
        return out


_REFT = None


def ref_tracking_lib():
    """oracle/_ref/libtracking_ref.so: Tracking::{ORB,LK}_Robust_Find_MuliImage_MatchedFeatures and deleteBadmatchFeatures compiled
    from the text of /root/reference/src/tracking.cpp:534-660 (cut out at build time), or None if unavailable."""
    global _REFT
    if _REFT is None:
        if build_ref() is None:
            return None
        so = os.path.join(_HERE, "_ref", "libtracking_ref.so")
        if not os.path.exists(so):
            return None
        L = C.CDLL(so)
        L.ref_orb_robust_match.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                           C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_lk_robust_track.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_double] + [C.c_void_p] * 4 + [C.c_int]
        _REFT = L
    return _REFT


def ref_orb_robust_match(last_kl, last_dl, last_kr, last_dr, cur_kl, cur_dl, fme=3.0):
    """The reference's compiled ORB association (tracking.cpp:534-581) -> (t2_left, t1_left, t1_right) point lists."""
    L = ref_tracking_lib()
    arrs = [np.ascontiguousarray(a) for a in (last_kl, last_dl, last_kr, last_dr, cur_kl, cur_dl)]
    cap = max(len(last_kl), 1)
    out = [np.zeros((cap, 2), np.float32) for _ in range(3)]
    n = L.ref_orb_robust_match(_p(arrs[0]), _p(arrs[1]), len(arrs[0]), _p(arrs[2]), _p(arrs[3]), len(arrs[2]), _p(arrs[4]), _p(arrs[5]),
                               len(arrs[4]), fme, _p(out[0]), _p(out[1]), _p(out[2]), cap)
    assert 0 <= n <= cap
    return tuple(o[:n] for o in out)


def ref_lk_robust_track(last_left, last_right, cur_left, cur_right, seeds, fme=3.0):
    """The reference's compiled 4-call LK chain + deleteBadmatchFeatures (tracking.cpp:583-660) ->
    (t2_left, t2_right, t1_left, t1_right) lists of the surviving tracks."""
    L = ref_tracking_lib()
    imgs = [_img(a) for a in (last_left, last_right, cur_left, cur_right)]
    seeds = np.ascontiguousarray(seeds, np.float32).reshape(-1, 2)
    cap = max(len(seeds), 1)
    out = [np.zeros((cap, 2), np.float32) for _ in range(4)]
    n = L.ref_lk_robust_track(_p(imgs[0]), _p(imgs[1]), _p(imgs[2]), _p(imgs[3]), imgs[0].shape[1], imgs[0].shape[0], _p(seeds), len(seeds), fme,
                              _p(out[0]), _p(out[1]), _p(out[2]), _p(out[3]), cap)
    assert 0 <= n <= cap
    return tuple(o[:n] for o in out)


def orb_match_filter(idx1, dist1, idx2, dist2, n_last_right, n_cur_left, last_left, last_right, fme=3.0):
    n = len(idx1)
    acc = np.zeros((max(n, 1), 3), np.int32)
    idx1, dist1, idx2, dist2 = (np.ascontiguousarray(a, np.int32) for a in (idx1, dist1, idx2, dist2))
    last_left = np.ascontiguousarray(last_left, KP_DTYPE)
    last_right = np.ascontiguousarray(last_right, KP_DTYPE)
    m = lib().orc_orb_match_filter(_p(idx1), _p(dist1), _p(idx2), _p(dist2), C.c_int(n), C.c_int(n_last_right), C.c_int(n_cur_left),
                                   _p(last_left), _p(last_right), C.c_double(fme), _p(acc))
    return acc[:m].copy()


def lk_chain_filter(p0, p1, p2, p3, p0r, s0, s1, s2, s3, fme=3.0):
    n = len(p0)
    keep = np.zeros(max(n, 1), np.uint8)
    f = [np.ascontiguousarray(a, np.float32).reshape(-1, 2) for a in (p0, p1, p2, p3, p0r)]
    s = [np.ascontiguousarray(a, np.uint8) for a in (s0, s1, s2, s3)]
    lib().orc_lk_chain_filter(C.c_int(n), *[_p(a) for a in f], *[_p(a) for a in s], C.c_double(fme), _p(keep))
    return keep[:n].astype(bool)


def triangulate_points(p1, p2, pts1, pts2):
    p1 = np.ascontiguousarray(p1, np.float64).reshape(12)
    p2 = np.ascontiguousarray(p2, np.float64).reshape(12)
    pts1 = np.ascontiguousarray(pts1, np.float32).reshape(-1, 2)
    pts2 = np.ascontiguousarray(pts2, np.float32).reshape(-1, 2)
    n = len(pts1)
    o4 = np.zeros((n, 4), np.float32)
    o3 = np.zeros((n, 3), np.float32)
    lib().orc_triangulate_points(_p(p1), _p(p2), _p(pts1), _p(pts2), C.c_int(n), _p(o4), _p(o3))
    return o4, o3


def frontend_sequence(left, right, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, mode=3, threads=1):
    left = np.ascontiguousarray(left, np.uint8)
    right = np.ascontiguousarray(right, np.uint8)
    f, h, w = left.shape
    return lib().orc_frontend_sequence(_p(left), _p(right), C.c_int(f), C.c_int(w), C.c_int(h), C.c_int(nfeatures),
                                       C.c_float(scale_factor), C.c_int(nlevels), C.c_int(ini_th), C.c_int(min_th), C.c_int(mode),
                                       C.c_int(threads))


# ---- pose from points (oracle_pnp.cpp) -------------------------------------------------------------------------------
def svd(a):
    """cv2.SVDecomp for a square matrix up to 12 x 12 -> (w, u, vt)."""
    a = np.ascontiguousarray(a, np.float64)
    n = a.shape[0]
    w, u, vt = np.zeros(n), np.zeros((n, n)), np.zeros((n, n))
    lib().orc_svd(_p(a), C.c_int(n), _p(w), _p(u), _p(vt))
    return w, u, vt


def epnp(K, p3, p2):
    """EPnP on all n points -> (ok, R 3x3, t 3)."""
    K = np.ascontiguousarray(K, np.float64).reshape(9)
    p3 = np.ascontiguousarray(p3, np.float32).reshape(-1, 3)
    p2 = np.ascontiguousarray(p2, np.float32).reshape(-1, 2)
    R = np.zeros(9)
    t = np.zeros(3)
    ok = lib().orc_epnp(_p(K), _p(p3), _p(p2), C.c_int(len(p3)), _p(R), _p(t))
    return bool(ok), R.reshape(3, 3), t


def pnp_refine(K, p3, p2, rvec, tvec, mask=None):
    """solvePnP(SOLVEPNP_ITERATIVE, useExtrinsicGuess=True) -> (rvec, tvec, LM iterations)."""
    K = np.ascontiguousarray(K, np.float64).reshape(9)
    p3 = np.ascontiguousarray(p3, np.float32).reshape(-1, 3)
    p2 = np.ascontiguousarray(p2, np.float32).reshape(-1, 2)
    param = np.concatenate([np.asarray(rvec, np.float64).reshape(3), np.asarray(tvec, np.float64).reshape(3)])
    m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
    it = lib().orc_pnp_refine(_p(K), _p(p3), _p(p2), C.c_int(len(p3)), None if m is None else _p(m), _p(param))
    return param[:3].copy(), param[3:].copy(), it


def pnp_ransac(K, p3, p2, iterations=500, reproj_err=0.5, confidence=0.99):
    """cv::solvePnPRansac as called at tracking.cpp:486-488 -> dict(ok, rvec, tvec, R, mask, n_inliers, best_iter, iters)."""
    K = np.ascontiguousarray(K, np.float64).reshape(9)
    p3 = np.ascontiguousarray(p3, np.float32).reshape(-1, 3)
    p2 = np.ascontiguousarray(p2, np.float32).reshape(-1, 2)
    n = len(p3)
    mask = np.zeros(max(n, 1), np.uint8)
    out = np.zeros(15)
    info = np.zeros(3, np.int32)
    ok = lib().orc_pnp_ransac(_p(K), _p(p3), _p(p2), C.c_int(n), C.c_int(iterations), C.c_float(reproj_err), C.c_double(confidence), _p(mask),
                              _p(out), _p(info))
    return dict(ok=bool(ok), rvec=out[:3].copy(), tvec=out[3:6].copy(), R=out[6:].reshape(3, 3).copy(), mask=mask[:n].astype(bool),
                n_inliers=int(info[0]), best_iter=int(info[1]), iters=int(info[2]))
