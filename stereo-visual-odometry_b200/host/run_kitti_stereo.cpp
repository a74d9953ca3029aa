// host/run_kitti_stereo.cpp -- the front-end of app/run_kitti_stereo.cpp (reference: app/run_kitti_stereo.cpp:5-19 -> System::Init /
// Run / Step, System.cpp:28-104) on the B200 path: a KITTI-layout dataset directory is decoded by the frame source
// (svo_source, System::NextFrame_kitti) and pushed through the throughput mode in batches; per frame it prints what
// Tracking::ORB_StereoF2F_PnP_Track / LK_StereoF2F_PnP_Track hand to the pose back-end (number of matched / tracked points).
// With --calib the pose step runs too (triangulation of (t1_left, t1_right) + solvePnPRansac against left_t, tracking.cpp:190-200 /
// 464-501, on the device) and the per-frame rvec / tvec / inlier counts are printed.  Pose gating, g2o and the Pangolin viewer are the
// reference's host code and are not part of this binary.
//
//   run_kitti_stereo <dataset_dir> [--mode orb|lk|both] [--batch 64] [--threads 8] [--width W --height H] [--nfeatures 1000] [--quiet]
//                    [--calib fx fy cx cy baseline_times_fx]     (KITTI 00: 718.856 718.856 607.1928 185.2157 386.1448)
//                    [--traj poses.txt]    with --calib: frame_pose_ after every frame (tracking.cpp:213-222 gates and accumulation, ORB list
//                                          when the ORB mode runs, else the LK list), 12 numbers per line like the KITTI odometry ground truth
// The image size is taken from the first left image unless given.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "svo_cuda.h"

static bool probe_size(const std::string& dir, int* w, int* h) {
    for (const char* ext : {"png", "pgm"}) {
        const std::string p = dir + "/image_0/000000." + ext;
        FILE* f = fopen(p.c_str(), "rb");
        if (!f) continue;
        unsigned char b[32] = {0};
        const size_t n = fread(b, 1, sizeof b, f);
        fclose(f);
        if (!strcmp(ext, "png") && n >= 24) {
            *w = (b[16] << 24) | (b[17] << 16) | (b[18] << 8) | b[19];
            *h = (b[20] << 24) | (b[21] << 16) | (b[22] << 8) | b[23];
            return true;
        }
        if (!strcmp(ext, "pgm") && sscanf((const char*)b, "P5 %d %d", w, h) == 2) return true;
    }
    return false;
}

int main(int argc, char** argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: %s <dataset_dir> [--mode orb|lk|both] [--batch N] [--threads N] [--width W --height H] [--nfeatures N] [--quiet]\n", argv[0]);
        return 2;
    }
    const std::string dir = argv[1];
    int mode = SVO_MODE_ORB | SVO_MODE_LK, batch = 64, threads = 8, w = 0, h = 0, nfeatures = 1000;
    bool quiet = false, have_calib = false;
    double calib[5] = {0, 0, 0, 0, 0};
    std::string traj_path;
    for (int i = 2; i < argc; ++i) {
        const std::string a = argv[i];
        auto val = [&]() { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "--mode") {
            const std::string m = val();
            mode = m == "orb" ? SVO_MODE_ORB : (m == "lk" ? SVO_MODE_LK : (SVO_MODE_ORB | SVO_MODE_LK));
        } else if (a == "--batch") batch = atoi(val());
        else if (a == "--threads") threads = atoi(val());
        else if (a == "--width") w = atoi(val());
        else if (a == "--height") h = atoi(val());
        else if (a == "--nfeatures") nfeatures = atoi(val());
        else if (a == "--quiet") quiet = true;
        else if (a == "--traj") traj_path = val();
        else if (a == "--calib") {
            for (double& v : calib) v = atof(val());
            have_calib = calib[0] > 0 && calib[1] > 0;
        }
    }
    if ((w <= 0 || h <= 0) && !probe_size(dir, &w, &h)) {
        fprintf(stderr, "cannot read %s/image_0/000000.png|pgm\n", dir.c_str());
        return 1;
    }
    svo_source* src = nullptr;
    if (svo_source_open(&src, dir.c_str(), 0, w, h, batch, 3, threads) != SVO_OK) {
        fprintf(stderr, "svo_source_open failed\n");
        return 1;
    }
    const int total = svo_source_frame_count(src);
    printf("dataset %s: %d stereo pairs of %dx%d, batch %d, %d decoder threads\n", dir.c_str(), total, w, h, batch, threads);
    svo_ctx* c = nullptr;
    if (svo_ctx_create(&c, 0, w, h, 2 * (2 * batch + 1)) != SVO_OK) {
        fprintf(stderr, "no CUDA device (there is no CPU fallback)\n");
        return 1;
    }
    auto die = [&](const char* what) {
        fprintf(stderr, "%s: %s\n", what, svo_last_error(c));
        exit(1);
    };
    if (svo_orb_configure(c, nfeatures, 1.2f, 8, 20, 7) != SVO_OK) die("svo_orb_configure");  // config/default.yaml ORBextractor.*
    if (svo_frontend_configure(c, w, h, mode, batch) != SVO_OK) die("svo_frontend_configure");
    if (have_calib) {  // Sensors::projMatr1_ / projMatr2_ (sensors.cpp:8-24) and the parameters of config/default.yaml:80-82
        const double P1[12] = {calib[0], 0, calib[2], 0, 0, calib[1], calib[3], 0, 0, 0, 1, 0};
        const double P2[12] = {calib[0], 0, calib[2], -calib[4], 0, calib[1], calib[3], 0, 0, 0, 1, 0};
        if (svo_frontend_set_projection(c, P1, P2) != SVO_OK) die("svo_frontend_set_projection");
        if (svo_frontend_set_pnp(c, 1, 500, 0.5f, (double)0.99f) != SVO_OK) die("svo_frontend_set_pnp");
    }
    std::vector<int32_t> n_orb(batch), n_lk(batch), n_left(batch), n_fast(batch), info_orb((size_t)4 * batch), info_lk((size_t)4 * batch);
    std::vector<double> pose_orb((size_t)16 * batch), pose_lk((size_t)16 * batch);
    svo_frontend_out out;
    memset(&out, 0, sizeof out);
    out.kp_cap = svo_orb_capacity(c, w, h);
    out.pt_cap = 1 << 20;
    out.n_orb_matched = n_orb.data();
    out.n_lk_matched = n_lk.data();
    out.n_left = n_left.data();
    out.n_fast = n_fast.data();
    if (have_calib) {
        out.orb_pose = pose_orb.data();
        out.orb_pnp_info = info_orb.data();
        out.lk_pose = pose_lk.data();
        out.lk_pnp_info = info_lk.data();
    }
    const auto t0 = std::chrono::steady_clock::now();
    const uint8_t *l = nullptr, *r = nullptr;
    int n = 0, frame = 0, pending = 0;
    long long sum_orb = 0, sum_lk = 0;
    // Tracking::frame_pose_ (tracking.h:117): accumulated on the host from the per-frame pose, with the reference's gates
    double frame_pose[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    int accepted = 0;
    FILE* traj = (!traj_path.empty() && have_calib) ? fopen(traj_path.c_str(), "w") : nullptr;
    auto accumulate = [&](const double* pose16, int ok, int inliers, int n_list, bool lk_mode) {
        if (!ok || n_list <= 0 || (double)inliers / n_list < 0.01) return false;  // inlier_rate (config/default.yaml:77)
        const double *R = pose16 + 6, *t = pose16 + 3;
        const float sy = (float)std::sqrt(R[0] * R[0] + R[3] * R[3]);  // rotationMatrixToEulerAngles (tracking.cpp:440-463)
        float ex, ey, ez;
        if (sy >= 1e-6f) {
            ex = (float)std::atan2(R[7], R[8]);
            ey = (float)std::atan2(-R[6], sy);
            ez = (float)std::atan2(R[3], R[0]);
        } else {
            ex = (float)std::atan2(-R[5], R[4]);
            ey = (float)std::atan2(-R[6], sy);
            ez = 0;
        }
        if (!(std::fabs(ex) < 0.1 && std::fabs(ey) < 0.1 && std::fabs(ez) < 0.1)) return false;
        const double n2 = t[0] * t[0] + t[1] * t[1] + t[2] * t[2];
        const double hi = lk_mode ? 100.0 : 10.0 * 10.0, lo = lk_mode ? 0.0005 * 0.0005 : 0.05 * 0.05;  // :312 literals / maxmove, minmove
        if (!(n2 < hi && n2 > lo)) return false;
        double inv[16] = {0}, prod[16];
        for (int r = 0; r < 3; ++r) {
            for (int c2 = 0; c2 < 3; ++c2) inv[4 * r + c2] = R[3 * c2 + r];
            inv[4 * r + 3] = -(R[r] * t[0] + R[3 + r] * t[1] + R[6 + r] * t[2]);
        }
        inv[15] = 1.0;
        for (int r = 0; r < 4; ++r)
            for (int c2 = 0; c2 < 4; ++c2) {
                double v = 0;
                for (int k2 = 0; k2 < 4; ++k2) v += frame_pose[4 * r + k2] * inv[4 * k2 + c2];
                prod[4 * r + c2] = v;
            }
        memcpy(frame_pose, prod, sizeof prod);
        return true;
    };
    auto report = [&](int k) {
        for (int i = 0; i < k; ++i, ++frame) {
            if (have_calib) {
                const bool use_orb = mode & SVO_MODE_ORB;
                const bool ok = use_orb ? accumulate(&pose_orb[(size_t)16 * i], info_orb[4 * i], info_orb[4 * i + 1], n_orb[i], false)
                                        : accumulate(&pose_lk[(size_t)16 * i], info_lk[4 * i], info_lk[4 * i + 1], n_lk[i], true);
                accepted += ok;
                if (traj) {
                    for (int q = 0; q < 12; ++q) fprintf(traj, "%.9e%c", frame_pose[q], q == 11 ? '\n' : ' ');
                }
            }
            sum_orb += n_orb[i];
            sum_lk += n_lk[i];
            if (!quiet)
                printf("frame %06d: keypoints %d, ORB matches %d, FAST corners %d, LK tracks %d\n", frame, (mode & SVO_MODE_ORB) ? n_left[i] : 0,
                       (mode & SVO_MODE_ORB) ? n_orb[i] : 0, (mode & SVO_MODE_LK) ? n_fast[i] : 0, (mode & SVO_MODE_LK) ? n_lk[i] : 0);
            if (!quiet && have_calib) {
                const double* po = &pose_orb[(size_t)16 * i];
                const double* pl = &pose_lk[(size_t)16 * i];
                if (mode & SVO_MODE_ORB)
                    printf("pose  %06d orb: ok %d inliers %d rvec %.6f %.6f %.6f tvec %.6f %.6f %.6f\n", frame, info_orb[4 * i], info_orb[4 * i + 1], po[0], po[1],
                           po[2], po[3], po[4], po[5]);
                if (mode & SVO_MODE_LK)
                    printf("pose  %06d lk : ok %d inliers %d rvec %.6f %.6f %.6f tvec %.6f %.6f %.6f\n", frame, info_lk[4 * i], info_lk[4 * i + 1], pl[0], pl[1],
                           pl[2], pl[3], pl[4], pl[5]);
            }
        }
    };
    // pipelined: run(k + 1) is issued before download(k); the source's ring slot of batch k is recycled by the next svo_source_next
    if (svo_source_next(src, &l, &r, &n) != SVO_OK) { fprintf(stderr, "%s\n", svo_source_error(src)); return 1; }
    while (n > 0) {
        if (svo_frontend_upload(c, l, r, n, (size_t)w) != SVO_OK) die("svo_frontend_upload");
        if (svo_frontend_run(c) != SVO_OK) die("svo_frontend_run");
        if (pending) {
            if (svo_frontend_download(c, &out) != SVO_OK) die("svo_frontend_download");
            report(pending);
        }
        pending = n;
        if (svo_source_next(src, &l, &r, &n) != SVO_OK) { fprintf(stderr, "%s\n", svo_source_error(src)); return 1; }
    }
    if (pending) {
        if (svo_frontend_download(c, &out) != SVO_OK) die("svo_frontend_download");
        report(pending);
    }
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("done: %d stereo pairs in %.3f s (%.1f pairs/s incl. decoding), ORB matches %lld, LK tracks %lld\n", frame, s, frame / s, sum_orb, sum_lk);
    if (have_calib)
        printf("trajectory: %d of %d frame-to-frame poses accepted, final position %.4f %.4f %.4f\n", accepted, frame > 0 ? frame - 1 : 0, frame_pose[3],
               frame_pose[7], frame_pose[11]);
    if (traj) fclose(traj);
    svo_source_close(src);
    svo_ctx_destroy(c);
    return frame == total ? 0 : 1;
}
