"""CPU suite: the C-ABI library loads, exports every symbol include/svo_cuda.h declares, and fails loudly without a GPU."""
import ctypes as C
import hashlib
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "svo_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(svo_[a-z0-9_]+)\s*\(", src)))


def test_exports_every_declared_symbol(pkg):
    L = pkg.lib()
    declared = _declared_symbols()
    assert len(declared) >= 25
    S = pkg.synth_lib()  # the host-only generator lives in its own library (no CUDA behind the CPU reference arm)
    missing = [s for s in declared if not hasattr(S if s in pkg.SYNTH_SYMBOLS else L, s)]
    assert not missing, missing
    assert sorted(pkg.ABI_SYMBOLS + pkg.SYNTH_SYMBOLS) == declared
    assert not any(hasattr(L, s) for s in pkg.SYNTH_SYMBOLS), "the generator must not be linked into the CUDA library"
    assert L.svo_abi_version() == 5


def test_keypoint_layout(pkg):
    assert pkg.KP_DTYPE.itemsize == 28 and pkg.KP_DTYPE.fields["octave"][1] == 20  # == cv::KeyPoint


def test_no_gpu_means_error_not_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = pkg.lib().svo_ctx_create(C.byref(h), 0, 640, 480, 2)
    assert rc == pkg.SVO_ERR_CUDA and not h.value
    with pytest.raises(pkg.SvoError):
        pkg.Context(0, 640, 480, 2)


def test_null_context_is_rejected(pkg):
    L = pkg.lib()
    assert L.svo_orb_configure(None, 1000, C.c_float(1.2), 8, 20, 7) == pkg.SVO_ERR_INVALID
    n = C.c_int(5)
    assert L.svo_orb_extract(None, None, 0, 0, C.c_size_t(0), None, None, 0, C.byref(n)) == pkg.SVO_ERR_INVALID


def test_synth_is_deterministic_and_stereo(pkg):
    l0, r0 = pkg.synth_stereo_frame(1000, 0, 1241, 376)
    l0b, _ = pkg.synth_stereo_frame(1000, 0, 1241, 376)
    assert np.array_equal(l0, l0b)
    assert hashlib.sha256(l0.tobytes()).hexdigest() == hashlib.sha256(l0b.tobytes()).hexdigest()
    l1, _ = pkg.synth_stereo_frame(1000, 1, 1241, 376)
    # 8-px disparity, 2-px/frame motion, up to the +-3 per-view sensor noise
    assert np.abs(l0[:, 8:].astype(int) - r0[:, :-8].astype(int)).max() <= 6
    assert np.abs(l0[:, 2:].astype(int) - l1[:, :-2].astype(int)).max() <= 6
    l16, _ = pkg.synth_stereo_frame(1000, 16, 1241, 376)  # canvas re-seeded every 16 frames
    assert np.abs(l0.astype(int) - l16.astype(int)).mean() > 10
    assert pkg.synth_lib().svo_synth_stereo_frame(1, 0, 0, 10, None, None, 0) == pkg.SVO_ERR_INVALID


def test_host_sources_typecheck_against_the_opencv_api():
    """INTEGRATION.md route A builds host/*.cpp with -DSVO_USE_REAL_OPENCV.  This image has no OpenCV C++, so the sources are
    type-checked (g++ -fsyntax-only) against tests/opencv_api_stub: declarations of the OpenCV classes they touch with the REAL
    proxy types (cv::_InputArray / _OutputArray / MatStep), not the Mat-reference shim the in-repo build uses."""
    for src in ("ORBextractor.cpp", "tracking_frontend.cpp"):
        __import__('subprocess').check_call(["g++", "-std=c++17", "-fsyntax-only", "-DSVO_USE_REAL_OPENCV", "-I" + os.path.join(ROOT, "tests", "opencv_api_stub"),
                               "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "stereo-visual-odometry_b200", "host", src)])
