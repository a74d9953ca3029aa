// tests/cpp/test_dropin.cpp -- drives the drop-in C++ classes (lzb_vio::ORBextractor, lzb_vio::Tracking front-end) exactly
// the way the reference's Tracking does and checks them against the CPU oracle.  Test code: it links the oracle.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../oracle/oracle.h"
#include "lzb_vio/tracking.h"
#include "svo_cuda.h"

#define CHECK(cond)                                                            \
    do {                                                                       \
        if (!(cond)) {                                                         \
            std::printf("CHECK FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            return 1;                                                          \
        }                                                                      \
    } while (0)

using namespace lzb_vio;

static Frame::Ptr make_frame(uint64_t seed, int t, int w, int h) {
    Frame::Ptr f = Frame::CreateFrame();
    f->left_img_.create(h, w, CV_8UC1);
    f->right_img_.create(h, w, CV_8UC1);
    svo_synth_stereo_frame(seed, t, w, h, f->left_img_.data, f->right_img_.data, f->left_img_.step);
    return f;
}

int main() {
    const int w = 640, h = 200, nfeat = 500;
    // ---- 1. ORBextractor::operator()
    {
        ORBextractor ex(nfeat, 1.2f, 8, 20, 7);
        CHECK(ex.GetLevels() == 8 && ex.GetScaleFactor() == 1.2f && ex.GetScaleFactors()[1] == 1.2f);
        Frame::Ptr f = make_frame(1000, 0, w, h);
        std::vector<cv::KeyPoint> kps;
        cv::Mat desc;
        ex(f->left_img_, cv::Mat(), kps, desc);
        orc::OrbResult r;
        orc::orb_extract(orc::OrbParams{nfeat, 1.2f, 8, 20, 7}, orc::Image{f->left_img_.data, w, h, f->left_img_.step}, r);
        CHECK(kps.size() == r.keypoints.size() && kps.size() >= (size_t)nfeat / 2);
        CHECK(std::memcmp(kps.data(), r.keypoints.data(), kps.size() * 28) == 0);  // bit-exact incl. angle and order
        CHECK(desc.rows == (int)kps.size() && desc.cols == 32);
        int bits = 0;
        for (int i = 0; i < desc.rows; ++i)
            for (int b = 0; b < 32; ++b) bits += __builtin_popcount(desc.ptr(i)[b] ^ r.descriptors[(size_t)i * 32 + b]);
        CHECK(bits <= 2);
        // mvImagePyramid: level pixels + 19-px reflect-101 border around the ROI
        CHECK((int)ex.mvImagePyramid.size() == 8);
        for (int l = 0; l < 8; ++l) {
            const cv::Mat& m = ex.mvImagePyramid[l];
            const orc::OwnedImage& o = r.pyramid[l];
            CHECK(m.cols == o.w && m.rows == o.h);
            for (int y = 0; y < o.h; ++y) CHECK(std::memcmp(m.ptr(y), &o.buf[(size_t)y * o.w], o.w) == 0);
            CHECK(m.ptr(0)[-1] == m.ptr(0)[1] && m.ptr(0)[-19] == m.ptr(0)[19] && m.ptr(-3)[5] == m.ptr(3)[5]);
            CHECK(m.ptr(o.h - 1 + 19)[o.w - 1 + 19] == m.ptr(o.h - 1 - 19)[o.w - 1 - 19]);
        }
        cv::Mat none;
        std::vector<cv::KeyPoint> k2 = kps;
        ex(none, cv::Mat(), k2, desc);  // empty image: silent return, outputs untouched
        CHECK(k2.size() == kps.size());
        cv::Mat flat(h, w, CV_8UC1);
        std::memset(flat.data, 90, (size_t)h * flat.step);
        ex(flat, cv::Mat(), k2, desc);
        CHECK(k2.empty() && desc.empty());
    }
    // ---- 2. Tracking, ORB mode (config/default.yaml track_mode)
    {
        Parameter::Ptr p(new Parameter);
        p->nFeatures_ = nfeat;
        p->track_mode_ = "ORB_stereof2f_pnp";
        Tracking tr(nullptr, p, nullptr);
        int calls = 0;
        size_t last_n = 0;
        tr.pose_backend_ = [&](const std::vector<cv::Point2f>& a, const std::vector<cv::Point2f>& b, const std::vector<cv::Point2f>& c) {
            ++calls;
            last_n = a.size();
            return a.size() == b.size() && b.size() == c.size();
        };
        Frame::Ptr f0 = make_frame(1000, 3, w, h), f1 = make_frame(1000, 4, w, h);
        CHECK(tr.AddFrame(f0) && tr.GetStatus() == TrackingStatus::TRACKING_GOOD && calls == 0);
        CHECK(tr.AddFrame(f1) && calls == 1 && last_n >= 50);
        // oracle on the descriptors the extractor produced
        const int nq = f0->left_Descriptors_.rows, nr = f0->right_Descriptors_.rows, nc = f1->left_Descriptors_.rows;
        std::vector<int32_t> i1(nq), d1(nq), i2(nq), d2(nq), acc((size_t)3 * nq + 3);
        orc::match_hamming(f0->left_Descriptors_.data, nq, f0->right_Descriptors_.data, nr, i1.data(), d1.data());
        orc::match_hamming(f0->left_Descriptors_.data, nq, f1->left_Descriptors_.data, nc, i2.data(), d2.data());
        CHECK((int)tr.last_match1_.size() == nq && (int)tr.last_match2_.size() == nq);
        for (int i = 0; i < nq; ++i) {
            CHECK(tr.last_match1_[i].trainIdx == i1[i] && tr.last_match1_[i].distance == (float)d1[i] && tr.last_match1_[i].queryIdx == i);
            CHECK(tr.last_match2_[i].trainIdx == i2[i] && tr.last_match2_[i].distance == (float)d2[i]);
        }
        std::vector<orc::KeyPoint> ll(nq), lr(nr);
        for (int i = 0; i < nq; ++i) std::memcpy(&ll[i], &f0->features_left_[i]->position_, 28);
        for (int i = 0; i < nr; ++i) std::memcpy(&lr[i], &f0->features_right_[i]->position_, 28);
        const int na = orc::orb_match_filter(i1.data(), d1.data(), i2.data(), d2.data(), nq, nr, nc, ll.data(), lr.data(), 3.0, acc.data());
        CHECK((size_t)na == last_n);
    }
    // ---- 2b. the pose step as ORB_StereoF2F_PnP_Track runs it (tracking.cpp:186-200): triangulate (t1_left, t1_right), then
    //          OpenCV_EstimatePose_PnP(projMatrl, matched_t2_left, points3D_t1, rotation, translation)
    {
        Parameter::Ptr p(new Parameter);
        p->nFeatures_ = nfeat;
        p->track_mode_ = "ORB_stereof2f_pnp";
        Tracking tr(nullptr, p, nullptr);
        cv::Mat projl(3, 4, CV_64F), projr(3, 4, CV_64F);
        const double P1[12] = {718.856, 0, 607.1928, 0, 0, 718.856, 185.2157, 0, 0, 0, 1, 0};
        for (int i = 0; i < 12; ++i) {
            projl.at<double>(i / 4, i % 4) = P1[i];
            projr.at<double>(i / 4, i % 4) = P1[i];
        }
        projr.at<double>(0, 3) = -386.1448;
        bool called = false, pose_ok = false;
        cv::Mat rotation, translation, points3D;
        std::vector<cv::Point2f> t1l, t1r, t2l;
        tr.pose_backend_ = [&](const std::vector<cv::Point2f>& a, const std::vector<cv::Point2f>& b, const std::vector<cv::Point2f>& c) {
            called = true;
            t1l = a; t1r = b; t2l = c;
            tr.TriangulateMatched(projl, projr, t1l, t1r, points3D);
            pose_ok = tr.OpenCV_EstimatePose_PnP(projl, t2l, points3D, rotation, translation);
            return pose_ok;
        };
        Frame::Ptr f0 = make_frame(1000, 3, w, h), f1 = make_frame(1000, 4, w, h);
        CHECK(tr.AddFrame(f0));
        tr.AddFrame(f1);
        CHECK(called && points3D.rows == (int)t2l.size() && rotation.rows == 3 && rotation.cols == 3 && translation.rows == 3);
        const int n = (int)t2l.size();
        std::vector<float> o4((size_t)4 * n), o3((size_t)3 * n);
        double P2[12];
        for (int i = 0; i < 12; ++i) P2[i] = P1[i];
        P2[3] = -386.1448;
        orc::triangulate_points(P1, P2, &t1l[0].x, &t1r[0].x, n, o4.data(), o3.data());
        const float* g3 = points3D.ptr<float>();
        for (int i = 0; i < 3 * n; ++i) CHECK(std::fabs(g3[i] - o3[i]) <= 2e-5f * std::fabs(o3[i]) + 1e-6f);
        // the oracle on the 3-D points the device produced (so that both sides see the same input)
        const double K[9] = {718.856, 0, 607.1928, 0, 718.856, 185.2157, 0, 0, 1};
        std::vector<uint8_t> mask(n);
        double out[15];
        int info[3];
        const int ok = orc::solve_pnp_ransac(K, g3, &t2l[0].x, n, 500, 0.5f, (double)0.99f, mask.data(), out, info);
        CHECK((int)tr.last_inliers_.size() == info[0]);
        size_t j = 0;
        for (int i = 0; i < n; ++i)
            if (mask[i]) CHECK(j < tr.last_inliers_.size() && tr.last_inliers_[j++] == i);
        for (int i = 0; i < 9; ++i) CHECK(std::fabs(rotation.at<double>(i / 3, i % 3) - out[6 + i]) <= 2e-6);  // LM stopping point, see tests/test_pnp_gpu.py
        for (int i = 0; i < 3; ++i) CHECK(std::fabs(translation.at<double>(i, 0) - out[3 + i]) <= 2e-6);
        CHECK(pose_ok == (ok && !((double)info[0] / n < 0.01)));
        // ---- 2c. the same frames with the projection matrices handed over and no pose_backend_: the reference's whole tail
        //          (tracking.cpp:186-246 minus the display) -> frame_pose_ = [R t; 0 1]^-1 when the gates pass
        Tracking tr2(nullptr, p, nullptr);
        tr2.SetProjection(projl, projr);
        CHECK(tr2.AddFrame(make_frame(1000, 3, w, h)));
        const bool accepted = tr2.AddFrame(make_frame(1000, 4, w, h));
        const cv::Vec3f e = tr.rotationMatrixToEulerAngles(rotation);
        const double tx = translation.at<double>(0, 0), ty = translation.at<double>(1, 0), tz = translation.at<double>(2, 0), n2 = tx * tx + ty * ty + tz * tz;
        const bool expect = pose_ok && std::fabs(e[0]) < 0.1 && std::fabs(e[1]) < 0.1 && std::fabs(e[2]) < 0.1 && n2 < 100.0 && n2 > 0.05 * 0.05;
        CHECK(accepted == expect);
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) {
                double want = r == c ? 1.0 : 0.0;
                if (expect && r < 3 && c < 3) want = rotation.at<double>(c, r);
                if (expect && r < 3 && c == 3)
                    want = -(rotation.at<double>(0, r) * tx + rotation.at<double>(1, r) * ty + rotation.at<double>(2, r) * tz);
                CHECK(std::fabs(tr2.frame_pose_.at<double>(r, c) - want) <= 1e-9);
            }
        std::printf("pose step: %d of %d inliers, |t| = %.4f, accepted %d\n", info[0], n, std::sqrt(n2), (int)accepted);
    }
    // ---- 3. Tracking, LK mode
    {
        Parameter::Ptr p(new Parameter);
        p->nFeatures_ = nfeat;
        p->track_mode_ = "LK_stereof2f_pnp";
        Tracking tr(nullptr, p, nullptr);
        std::vector<cv::Point2f> g1l, g1r, g2l;
        tr.pose_backend_ = [&](const std::vector<cv::Point2f>& a, const std::vector<cv::Point2f>& b, const std::vector<cv::Point2f>& c) {
            g1l = a; g1r = b; g2l = c;
            return true;
        };
        Frame::Ptr f0 = make_frame(1000, 5, w, h), f1 = make_frame(1000, 6, w, h);
        CHECK(tr.AddFrame(f0));
        CHECK(tr.AddFrame(f1));
        std::vector<orc::KeyPoint> fast;
        orc::fast9_detect(orc::Image{f0->left_img_.data, w, h, f0->left_img_.step}, 20, true, fast);
        CHECK(fast.size() == f0->features_left_.size());
        const int n = (int)fast.size();
        std::vector<float> p0(2 * n), p1(2 * n), p2(2 * n), p3(2 * n), p0r(2 * n), err(n);
        std::vector<uint8_t> s0(n), s1(n), s2(n), s3(n), keep(n);
        for (int i = 0; i < n; ++i) { p0[2 * i] = fast[i].x; p0[2 * i + 1] = fast[i].y; }
        orc::Image L0{f0->left_img_.data, w, h, f0->left_img_.step}, R0{f0->right_img_.data, w, h, f0->right_img_.step};
        orc::Image L1{f1->left_img_.data, w, h, f1->left_img_.step}, R1{f1->right_img_.data, w, h, f1->right_img_.step};
        orc::calc_optical_flow_pyr_lk(L0, R0, p0.data(), n, p1.data(), s0.data(), err.data(), 21, 3, 30, 0.01, 0.001);
        orc::calc_optical_flow_pyr_lk(R0, R1, p1.data(), n, p2.data(), s1.data(), err.data(), 21, 3, 30, 0.01, 0.001);
        orc::calc_optical_flow_pyr_lk(R1, L1, p2.data(), n, p3.data(), s2.data(), err.data(), 21, 3, 30, 0.01, 0.001);
        orc::calc_optical_flow_pyr_lk(L1, L0, p3.data(), n, p0r.data(), s3.data(), err.data(), 21, 3, 30, 0.01, 0.001);
        const int kept = orc::lk_chain_filter(n, p0.data(), p1.data(), p2.data(), p3.data(), p0r.data(), s0.data(), s1.data(), s2.data(), s3.data(), 3.0, keep.data());
        CHECK((int)g1l.size() == kept && kept >= 50);
        int j = 0;
        for (int i = 0; i < n; ++i) {
            if (!keep[i]) continue;
            CHECK(g1l[j].x == p0[2 * i] && g1l[j].y == p0[2 * i + 1]);
            CHECK(std::fabs(g1r[j].x - p1[2 * i]) <= 0.01f && std::fabs(g1r[j].y - p1[2 * i + 1]) <= 0.01f);
            CHECK(std::fabs(g2l[j].x - p3[2 * i]) <= 0.01f && std::fabs(g2l[j].y - p3[2 * i + 1]) <= 0.01f);
            ++j;
        }
    }
    std::printf("DROPIN PASS\n");
    return 0;
}
