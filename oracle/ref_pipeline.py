"""CPU reference arm of bench.py (TEST / BASELINE INFRASTRUCTURE, never on the product path).

The whole reference binary cannot be built in this image (no OpenCV C++ / Sophus / glog / g2o / Pangolin),
so the reference's CPU front-end is reproduced as closely as the image allows:
  * the OpenCV primitives the reference calls directly from tracking.cpp -- cv::FAST (tracking.cpp:101), BFMatcher
    (tracking.cpp:539-544), cv::calcOpticalFlowPyrLK (tracking.cpp:593-618) -- are the REAL OpenCV ones (cv2 4.13, SIMD,
    one OpenCV thread per worker);
  * ORBextractor::operator() is the reference's OWN src/ORBextractor.cpp, compiled where it lies against oracle/ref_shim into
    oracle/_ref/liborb_ref.so (system malloc, the reference's -O3 flags; the OpenCV calls inside it are the oracle's cv2-pinned
    restatements).  Where that library is absent, the C++ oracle port (oracle_orb.cpp, -O3 AVX2) stands in -- same speed
    within a few percent, same results up to the octree's heap-address tie-break;
  * the reference is single-threaded; to give it every host core the sequence is split over one worker process per core
    (contiguous chunks overlapping by one frame, the same sharding rule as the GPU path).
"""
import ctypes as C
import math
import multiprocessing as mp
import os
import sys
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_G = {}


class Cv2PrimitiveOrb:
    """ORBextractor::operator() with the REAL OpenCV image primitives (variant 2): cv2.resize per level
    (ORBextractor.cpp:1074), cv2.FAST on every cell window with the 20 -> 7 fallback (:742-781; ROI views, no copies),
    cv2.GaussianBlur per level (:1035); the reference's own loops around them (octree, IC_Angle, steered BRIEF) run compiled in
    liboracle.so (orc_orb_finish).  Results equal the oracle's (tests/test_oracle_cpu.py)."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        import cv2
        sys.path.insert(0, _HERE)
        import oracle_py as O
        self.cv2, self.O = cv2, O
        self.p = (nfeatures, scale_factor, nlevels, ini_th, min_th)
        self.t = O.orb_tables(nfeatures, scale_factor, nlevels, ini_th, min_th)
        self.fast_ini = cv2.FastFeatureDetector_create(ini_th, True)
        self.fast_min = cv2.FastFeatureDetector_create(min_th, True)

    def extract(self, img):
        cv2, O = self.cv2, self.O
        nfeatures, scale_factor, nlevels, ini_th, min_th = self.p
        h0, w0 = img.shape
        levels, blurred, cands = [np.ascontiguousarray(img)], [], []
        for l in range(1, nlevels):
            s = self.t[1][l]  # inverse scale factors (orb_tables: scale, inv_scale, features per level, umax)
            w, h = int(np.rint(np.float32(w0) * s)), int(np.rint(np.float32(h0) * s))
            levels.append(cv2.resize(levels[l - 1], (w, h), interpolation=cv2.INTER_LINEAR))
        EDGE = 19
        for l in range(nlevels):
            lvl = levels[l]
            h, w = lvl.shape
            minBX = minBY = EDGE - 3
            maxBX, maxBY = w - EDGE + 3, h - EDGE + 3
            pts = []
            if maxBX - minBX >= 30 and maxBY - minBY >= 30:
                width, height = float(maxBX - minBX), float(maxBY - minBY)
                nCols, nRows = int(width / 30), int(height / 30)
                wCell, hCell = int(math.ceil(width / nCols)), int(math.ceil(height / nRows))
                for i in range(nRows):
                    iniY = minBY + i * hCell
                    maxY = iniY + hCell + 6
                    if iniY >= maxBY - 3:
                        continue
                    maxY = min(maxY, maxBY)
                    for j in range(nCols):
                        iniX = minBX + j * wCell
                        maxX = iniX + wCell + 6
                        if iniX >= maxBX - 6:
                            continue
                        maxX = min(maxX, maxBX)
                        roi = lvl[iniY:maxY, iniX:maxX]
                        kps = self.fast_ini.detect(roi)
                        if not kps:
                            kps = self.fast_min.detect(roi)
                        for k in kps:
                            pts.append((k.pt[0] + j * wCell, k.pt[1] + i * hCell, k.response))
            c = np.zeros(len(pts), O.KP_DTYPE)
            if pts:
                a = np.asarray(pts, np.float32)
                c["x"], c["y"], c["response"] = a[:, 0], a[:, 1], a[:, 2]
                c["size"], c["angle"], c["class_id"] = 7.0, -1.0, -1
            cands.append(c)
            blurred.append(cv2.GaussianBlur(lvl, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101))
        cap = nfeatures + 64 * nlevels + 64
        kps = np.zeros(cap, O.KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        lv_p = (C.c_void_p * nlevels)(*[a.ctypes.data for a in levels])
        bl_p = (C.c_void_p * nlevels)(*[a.ctypes.data for a in blurred])
        ws = (C.c_int * nlevels)(*[a.shape[1] for a in levels])
        hs = (C.c_int * nlevels)(*[a.shape[0] for a in levels])
        nc = (C.c_int * nlevels)(*[len(c) for c in cands])
        allc = np.ascontiguousarray(np.concatenate(cands)) if sum(len(c) for c in cands) else np.zeros(1, O.KP_DTYPE)
        n = O.lib().orc_orb_finish(C.c_int(nfeatures), C.c_float(scale_factor), C.c_int(nlevels), C.c_int(ini_th), C.c_int(min_th), lv_p, bl_p,
                                   ws, hs, allc.ctypes.data_as(C.c_void_p), nc, kps.ctypes.data_as(C.c_void_p),
                                   desc.ctypes.data_as(C.c_void_p), C.c_int(cap))
        assert n <= cap
        return kps[:n].copy(), desc[:n].copy()


def _worker(args):
    u0, u1, mode, nfeat, variant, cv_threads = args
    import cv2
    cv2.setNumThreads(cv_threads)
    sys.path.insert(0, _HERE)
    import oracle_py as O
    left, right = _G["left"], _G["right"]
    if variant == "cv2prim":
        orb = Cv2PrimitiveOrb(nfeat, 1.2, 8, 20, 7)
    elif variant == "port":
        orb = O.OrbOracle(nfeat, 1.2, 8, 20, 7)
    else:
        orb = O.RefOrb(nfeat, 1.2, 8, 20, 7, monotone=False) if O.ref_lib() is not None else O.OrbOracle(nfeat, 1.2, 8, 20, 7)
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    fast = cv2.FastFeatureDetector_create(20, True)
    lk = dict(winSize=(21, 21), maxLevel=3, criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01), flags=0,
              minEigThreshold=1e-3)
    checksum = 0
    last = None
    P1 = np.array([[718.856, 0, 607.1928, 0], [0, 718.856, 185.2157, 0], [0, 0, 1, 0]], np.float64)   # KITTI sequence 00
    P2 = np.array([[718.856, 0, 607.1928, -386.1448], [0, 718.856, 185.2157, 0], [0, 0, 1, 0]], np.float64)

    def pose_step(t1_left, t1_right, t2_left):
        """tracking.cpp:190-200 / 464-501 with the real OpenCV calls: triangulatePoints, convertPointsFromHomogeneous, solvePnPRansac."""
        if len(t2_left) < 5:
            return 0
        p4 = cv2.triangulatePoints(P1, P2, np.ascontiguousarray(t1_left.T), np.ascontiguousarray(t1_right.T))
        p3 = cv2.convertPointsFromHomogeneous(p4.T)
        ok, rvec, tvec, inl = cv2.solvePnPRansac(p3, np.ascontiguousarray(t2_left), P1[:, :3].copy(), np.zeros((4, 1)), rvec=np.zeros((3, 1)),
                                                 tvec=np.zeros((3, 1)), useExtrinsicGuess=True, iterationsCount=500, reprojectionError=0.5,
                                                 confidence=float(np.float32(0.99)), flags=cv2.SOLVEPNP_ITERATIVE)
        cv2.Rodrigues(rvec)
        return 0 if inl is None else len(inl)

    for t in range(u0, u1 + 1):  # frame u0 only warms up the "last frame"
        cur = {}
        L, R = left[t], right[t]
        if mode & 1:
            cur["kl"], cur["dl"] = orb.extract(L)
            cur["kr"], cur["dr"] = orb.extract(R)
            if last is not None and len(last["dl"]) and len(last["dr"]) and len(cur["dl"]):
                m1 = bf.match(last["dl"], last["dr"])
                m2 = bf.match(last["dl"], cur["dl"])
                i1 = np.fromiter((m.trainIdx for m in m1), np.int32, len(m1))
                d1 = np.fromiter((m.distance for m in m1), np.int32, len(m1))
                i2 = np.fromiter((m.trainIdx for m in m2), np.int32, len(m2))
                d2 = np.fromiter((m.distance for m in m2), np.int32, len(m2))
                acc = O.orb_match_filter(i1, d1, i2, d2, len(last["dr"]), len(cur["dl"]), last["kl"], last["kr"], 3.0)
                checksum += len(acc)
                if (mode & 4) and len(acc):
                    a = np.asarray(acc).reshape(-1, 3)
                    xy = lambda k, idx: np.stack([k["x"][idx], k["y"][idx]], 1).astype(np.float32)  # noqa: E731
                    checksum += pose_step(xy(last["kl"], a[:, 0]), xy(last["kr"], a[:, 1]), xy(cur["kl"], a[:, 2]))
        if mode & 2:
            kps = fast.detect(L)
            cur["pts"] = np.array([k.pt for k in kps], np.float32).reshape(-1, 1, 2)
            if last is not None and len(last["pts"]):
                p0 = last["pts"]
                p1, s1, _ = cv2.calcOpticalFlowPyrLK(last["L"], last["R"], p0, None, **lk)
                p2, s2, _ = cv2.calcOpticalFlowPyrLK(last["R"], R, p1, None, **lk)
                p3, s3, _ = cv2.calcOpticalFlowPyrLK(R, L, p2, None, **lk)
                p0r, s4, _ = cv2.calcOpticalFlowPyrLK(L, last["L"], p3, None, **lk)
                keep = O.lk_chain_filter(p0.reshape(-1, 2), p1.reshape(-1, 2), p2.reshape(-1, 2), p3.reshape(-1, 2), p0r.reshape(-1, 2),
                                         s1.ravel(), s2.ravel(), s3.ravel(), s4.ravel(), 3.0)
                checksum += int(keep.sum())
                if (mode & 4) and keep.any():
                    checksum += pose_step(p0.reshape(-1, 2)[keep], p1.reshape(-1, 2)[keep], p3.reshape(-1, 2)[keep])
            cur["L"], cur["R"] = L, R
        last = cur
    return checksum


def orb_kind():
    """'reference' when ORBextractor runs from the reference's compiled source (oracle/_ref), else 'port'."""
    sys.path.insert(0, _HERE)
    import oracle_py as O
    return "reference" if O.ref_lib() is not None else "port"


class ReferencePool:
    """One worker process per host core over a fixed synthetic sequence; run() processes all units once."""

    def __init__(self, left, right, mode=3, nfeatures=1000, workers=None, variant="ref", cv_threads=1):
        """variant: 'ref' (ORBextractor from the reference's compiled source, falls back to the port), 'port' (C++ oracle port),
        'cv2prim' (real-OpenCV image primitives, Cv2PrimitiveOrb); cv_threads: OpenCV-internal threads per worker."""
        self.workers = workers or len(os.sched_getaffinity(0))
        _G["left"], _G["right"] = left, right
        self.units = left.shape[0] - 1
        self.workers = max(1, min(self.workers, self.units))
        self.mode, self.nfeat = mode, nfeatures
        self.pool = mp.get_context("fork").Pool(self.workers)
        self.chunks = []
        for k in range(self.workers):
            u0, u1 = self.units * k // self.workers, self.units * (k + 1) // self.workers
            if u1 > u0:
                self.chunks.append((u0, u1, mode, nfeatures, variant, cv_threads))

    def run(self):
        t0 = time.perf_counter()
        cs = sum(self.pool.map(_worker, self.chunks, chunksize=1))
        return time.perf_counter() - t0, cs

    def close(self):
        self.pool.close()
        self.pool.join()
