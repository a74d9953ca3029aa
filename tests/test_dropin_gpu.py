"""GPU: the drop-in C++ classes (lzb_vio::ORBextractor / Tracking front-end) driven like the reference drives them."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_dropin(pkg, orc, tmp_path):
    orc.lib()  # builds oracle/liboracle.so if needed
    pkgdir = os.path.join(ROOT, "stereo-visual-odometry_b200")
    if not os.path.exists(os.path.join(pkgdir, "liblzb_vio_b200.so")):
        subprocess.check_call(["make", "-C", os.path.join(pkgdir, "host"), "-s"])
    exe = str(tmp_path / "test_dropin")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_dropin.cpp"),
                           "-o", exe, "-L" + pkgdir, "-llzb_vio_b200", "-lsvo_b200", "-lsvo_synth", "-L" + os.path.join(ROOT, "oracle"), "-loracle",
                           "-Wl,-rpath," + pkgdir, "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-pthread"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "DROPIN PASS" in out.stdout, out.stdout + out.stderr
