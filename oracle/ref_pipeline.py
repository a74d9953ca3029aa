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
import multiprocessing as mp
import os
import sys
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_G = {}


def _worker(args):
    u0, u1, mode, nfeat = args
    import cv2
    cv2.setNumThreads(1)
    sys.path.insert(0, _HERE)
    import oracle_py as O
    left, right = _G["left"], _G["right"]
    orb = O.RefOrb(nfeat, 1.2, 8, 20, 7, monotone=False) if O.ref_lib() is not None else O.OrbOracle(nfeat, 1.2, 8, 20, 7)
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    fast = cv2.FastFeatureDetector_create(20, True)
    lk = dict(winSize=(21, 21), maxLevel=3, criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01), flags=0,
              minEigThreshold=1e-3)
    checksum = 0
    last = None
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

    def __init__(self, left, right, mode=3, nfeatures=1000, workers=None):
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
                self.chunks.append((u0, u1, mode, nfeatures))

    def run(self):
        t0 = time.perf_counter()
        cs = sum(self.pool.map(_worker, self.chunks, chunksize=1))
        return time.perf_counter() - t0, cs

    def close(self):
        self.pool.close()
        self.pool.join()
