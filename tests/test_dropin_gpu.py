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


def test_run_kitti_stereo_app(pkg, tmp_path):
    """The front-end of app/run_kitti_stereo.cpp: dataset directory -> svo_source -> throughput mode, as a plain C-ABI binary."""
    import numpy as np
    pkgdir = os.path.join(ROOT, "stereo-visual-odometry_b200")
    exe = os.path.join(pkgdir, "run_kitti_stereo")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.join(pkgdir, "host"), "-s"])
    w, h, F = 640, 200, 9
    left, right = pkg.synth_sequence(31, 0, F, w, h)
    pkg.write_sequence(str(tmp_path), left, right, fmt="png")
    out = subprocess.run([exe, str(tmp_path), "--mode", "both", "--batch", "4", "--threads", "3", "--nfeatures", "400", "--calib", "718.856", "718.856",
                          "607.1928", "185.2157", "386.1448", "--traj", str(tmp_path / "traj.txt")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = [l for l in out.stdout.splitlines() if l.startswith("frame ")]
    assert len(lines) == F and "done: 9 stereo pairs" in out.stdout
    # the same numbers through the Python binding
    ctx = pkg.Context(0, w, h, 2 * (F + 1))
    ctx.orb_configure(400, 1.2, 8, 20, 7)
    ctx.frontend_configure(w, h, pkg.MODE_ORB | pkg.MODE_LK, F)
    P1 = [718.856, 0, 607.1928, 0, 0, 718.856, 185.2157, 0, 0, 0, 1, 0]
    P2 = [718.856, 0, 607.1928, -386.1448, 0, 718.856, 185.2157, 0, 0, 0, 1, 0]
    ctx.frontend_set_projection(P1, P2)
    ctx.frontend_set_pnp(True, 500, 0.5)
    o = pkg.FrontendBuffers(F, ctx.orb_capacity(w, h), max(1024, w * h // 64),
                            fields=("n_left", "n_orb_matched", "n_fast", "n_lk_matched", "orb_pose", "orb_pnp_info", "lk_pose", "lk_pnp_info"))
    ctx.frontend_batch(left, right, o)
    for i, l in enumerate(lines):
        nums = [int(t.strip(",")) for t in l.replace(":", " ").split() if t.strip(",").isdigit()]
        assert nums[1:] == [int(o.n_left[i]), int(o.n_orb_matched[i]), int(o.n_fast[i]), int(o.n_lk_matched[i])], l
    # the pose lines (--calib): the app's batches of 4 and this single batch of 9 give the same per-frame poses
    poses = [l.split() for l in out.stdout.splitlines() if l.startswith("pose ")]
    assert len(poses) == 2 * F
    for p in poses:
        i, kind = int(p[1]), p[2].rstrip(":")
        info, pose = (o.orb_pnp_info, o.orb_pose) if kind == "orb" else (o.lk_pnp_info, o.lk_pose)
        vals = [t for t in p if t.replace(".", "").replace("-", "").isdigit()]
        assert int(p[p.index("ok") + 1]) == int(info[i, 0]) and int(p[p.index("inliers") + 1]) == int(info[i, 1]), p
        got = np.array([float(x) for x in p[p.index("rvec") + 1:p.index("rvec") + 4]] + [float(x) for x in p[p.index("tvec") + 1:p.index("tvec") + 4]])
        assert np.abs(got - pose[i, :6]).max() <= 1e-6, (p, pose[i, :6])
    # --traj: one 3 x 4 pose per frame, the first the identity, every later one the previous times an accepted inverse step or unchanged
    traj = np.loadtxt(str(tmp_path / "traj.txt")).reshape(-1, 3, 4)
    assert traj.shape[0] == F and np.allclose(traj[0], np.eye(4)[:3])
    acc = np.eye(4)
    for i in range(1, F):
        if o.orb_pnp_info[i, 0] and o.n_orb_matched[i] > 0 and o.orb_pnp_info[i, 1] / o.n_orb_matched[i] >= 0.01:
            R, t = o.orb_pose[i, 6:15].reshape(3, 3), o.orb_pose[i, 3:6]
            sy = np.float32(np.sqrt(R[0, 0] ** 2 + R[1, 0] ** 2))
            e = (np.arctan2(R[2, 1], R[2, 2]), np.arctan2(-R[2, 0], sy), np.arctan2(R[1, 0], R[0, 0]))
            if all(abs(np.float32(v)) < 0.1 for v in e) and 0.05 ** 2 < t @ t < 100:
                T = np.eye(4)
                T[:3, :3], T[:3, 3] = R.T, -R.T @ t
                acc = acc @ T
        assert np.abs(traj[i] - acc[:3]).max() <= 1e-7, i
    assert "trajectory:" in out.stdout
    ctx.close()
