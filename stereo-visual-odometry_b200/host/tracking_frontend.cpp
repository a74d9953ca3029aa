// host/tracking_frontend.cpp -- the front-end half of lzb_vio::Tracking over the C ABI.
// Replaces the OpenCV calls at tracking.cpp:101 (cv::FAST), :508-509 (ORBextractor), :539-544 (BFMatcher), :593-618
// (calcOpticalFlowPyrLK), :190-192 (triangulatePoints + convertPointsFromHomogeneous), :486-489 (solvePnPRansac + Rodrigues);
// the filters around them (:546-577, :623-660) are restated on the host as in the reference.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <stdexcept>

#include "lzb_vio/tracking.h"
#include "svo_cuda.h"

namespace lzb_vio {

Tracking::Tracking(System* system, Parameter::Ptr parameter, std::shared_ptr<Sensors> sensors)
    : system_(system), parameter_(parameter), sensors_(sensors) {
    Readparameter();
    mpORBextractorLeft_ = new ORBextractor(nFeatures_, fScaleFactor_, nLevels_, fIniThFAST_, fMinThFAST_);  // tracking.cpp:20
    mpORBextractorLeft_->SetExportPyramid(false);  // nothing in Tracking reads mvImagePyramid
    frame_pose_.create(4, 4, CV_64F);
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) frame_pose_.at<double>(r, c) = r == c ? 1.0 : 0.0;
}

Tracking::~Tracking() { delete mpORBextractorLeft_; }

void Tracking::Readparameter() {  // tracking.cpp:22-47, front-end fields
    feature_match_error_ = parameter_->feature_match_error_;
    track_mode_ = parameter_->track_mode_;
    num_features_tracking_ = parameter_->num_features_tracking_;
    inlier_rate_ = parameter_->inlier_rate_;
    iterationsCount_ = parameter_->iterationsCount_;
    reprojectionError_ = parameter_->reprojectionError_;
    confidence_ = parameter_->confidence_;
    minmove_ = parameter_->minmove_;
    maxmove_ = parameter_->maxmove_;
    nFeatures_ = parameter_->nFeatures_;
    fScaleFactor_ = parameter_->fScaleFactor_;
    nLevels_ = parameter_->nLevels_;
    fIniThFAST_ = parameter_->fIniThFAST_;
    fMinThFAST_ = parameter_->fMinThFAST_;
}

bool Tracking::AddFrame(Frame::Ptr frame) {  // tracking.cpp:49-77
    current_frame_ = frame;
    switch (status_) {
        case TrackingStatus::INITING:
            StereoInit_f2f();
            break;
        case TrackingStatus::TRACKING_GOOD: {
            const bool ok = Track();
            last_frame_ = current_frame_;
            return ok;
        }
        case TrackingStatus::LOST:
            break;
    }
    return true;
}

bool Tracking::StereoInit_f2f() {  // tracking.cpp:78-92
    if (track_mode_ == "ORB_stereof2f_pnp")
        Detect_MyORBFeatures();
    else
        Detect_OpenCVFASTFeatures();
    last_frame_ = current_frame_;
    status_ = TrackingStatus::TRACKING_GOOD;
    return true;
}

bool Tracking::Track() {  // tracking.cpp:115-128
    if (track_mode_ == "LK_stereof2f_pnp") return LK_StereoF2F_PnP_Track();
    if (track_mode_ == "ORB_stereof2f_pnp") return ORB_StereoF2F_PnP_Track();
    return false;
}

bool Tracking::ORB_StereoF2F_PnP_Track() {  // tracking.cpp:168-186; the pose back-end (:186-249) is pose_backend_
    Detect_MyORBFeatures();
    std::vector<cv::Point2f> matched_t1_left, matched_t2_left, matched_t1_right;
    const int n = ORB_Robust_Find_MuliImage_MatchedFeatures(matched_t2_left, matched_t1_left, matched_t1_right);
    if (n < num_features_tracking_) return false;
    if (pose_backend_) return pose_backend_(matched_t1_left, matched_t1_right, matched_t2_left);
    return projMatrl_.empty() ? true : EstimateAndAccumulatePose(matched_t1_left, matched_t1_right, matched_t2_left, false);
}

void Tracking::SetProjection(const cv::Mat& projMatrl, const cv::Mat& projMatrr) {
    projMatrl_ = projMatrl.clone();
    projMatrr_ = projMatrr.clone();
}

cv::Vec3f Tracking::rotationMatrixToEulerAngles(cv::Mat& R) {  // tracking.cpp:440-463
    const float sy = (float)std::sqrt(R.at<double>(0, 0) * R.at<double>(0, 0) + R.at<double>(1, 0) * R.at<double>(1, 0));
    const bool singular = sy < 1e-6;
    float x, y, z;
    if (!singular) {
        x = (float)std::atan2(R.at<double>(2, 1), R.at<double>(2, 2));
        y = (float)std::atan2(-R.at<double>(2, 0), sy);
        z = (float)std::atan2(R.at<double>(1, 0), R.at<double>(0, 0));
    } else {
        x = (float)std::atan2(-R.at<double>(1, 2), R.at<double>(1, 1));
        y = (float)std::atan2(-R.at<double>(2, 0), sy);
        z = 0;
    }
    return cv::Vec3f(x, y, z);
}

// tracking.cpp:186-246 (ORB mode) / 288-341 (LK mode) without the display calls
bool Tracking::EstimateAndAccumulatePose(std::vector<cv::Point2f>& matched_t1_left, std::vector<cv::Point2f>& matched_t1_right,
                                         std::vector<cv::Point2f>& matched_t2_left, bool lk_mode) {
    cv::Mat points3D_t1, rotation, translation;
    TriangulateMatched(projMatrl_, projMatrr_, matched_t1_left, matched_t1_right, points3D_t1);
    if (!OpenCV_EstimatePose_PnP(projMatrl_, matched_t2_left, points3D_t1, rotation, translation)) return false;
    const cv::Vec3f e = rotationMatrixToEulerAngles(rotation);
    if (!(std::fabs(e[1]) < 0.1 && std::fabs(e[0]) < 0.1 && std::fabs(e[2]) < 0.1)) return false;
    const double t[3] = {translation.at<double>(0, 0), translation.at<double>(1, 0), translation.at<double>(2, 0)};
    const double pose_norm2 = t[0] * t[0] + t[1] * t[1] + t[2] * t[2];
    // the LK branch compares against the literals 100 and 0.0005^2 (:312), the ORB branch against maxmove_ / minmove_ (:213)
    const double hi = lk_mode ? 100.0 : maxmove_ * maxmove_, lo = lk_mode ? 0.0005 * 0.0005 : minmove_ * minmove_;
    if (!(pose_norm2 < hi && pose_norm2 > lo)) return false;
    // rigid_body_transformation = [R t; 0 0 0 1]^-1 = [R^T, -R^T t]; frame_pose_ = frame_pose_ * rigid_body_transformation
    double inv[16] = {0};
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) inv[4 * r + c] = rotation.at<double>(c, r);
        inv[4 * r + 3] = -(rotation.at<double>(0, r) * t[0] + rotation.at<double>(1, r) * t[1] + rotation.at<double>(2, r) * t[2]);
    }
    inv[15] = 1.0;
    double prod[16];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += frame_pose_.at<double>(r, k) * inv[4 * k + c];
            prod[4 * r + c] = s;
        }
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) frame_pose_.at<double>(r, c) = prod[4 * r + c];
    Px_ = frame_pose_.at<double>(0, 3);
    Py_ = frame_pose_.at<double>(1, 3);
    Pz_ = frame_pose_.at<double>(2, 3);
    return true;
}

bool Tracking::LK_StereoF2F_PnP_Track() {  // tracking.cpp:258-288
    const int num_features_left = Detect_OpenCVFASTFeatures();
    if (num_features_left < 30) return false;
    std::vector<cv::Point2f> matched_t2_left, matched_t1_left, matched_t2_right, matched_t1_right;
    for (auto& feature : last_frame_->features_left_) matched_t1_left.push_back(feature->position_.pt);
    const int n = LK_Robust_Find_MuliImage_MatchedFeatures(matched_t2_left, matched_t2_right, matched_t1_left, matched_t1_right);
    if (n < num_features_tracking_) return false;
    if (pose_backend_) return pose_backend_(matched_t1_left, matched_t1_right, matched_t2_left);
    return projMatrl_.empty() ? true : EstimateAndAccumulatePose(matched_t1_left, matched_t1_right, matched_t2_left, true);
}

void Tracking::TriangulateMatched(cv::Mat& projMatrl, cv::Mat& projMatrr, std::vector<cv::Point2f>& points_t1_left,
                                  std::vector<cv::Point2f>& points_t1_right, cv::Mat& points3D_t1) {  // tracking.cpp:190-192
    const int n = (int)std::min(points_t1_left.size(), points_t1_right.size());
    points3D_t1.create(n, 1, CV_32FC3);
    if (n == 0) return;
    double p1[12], p2[12];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c) {
            p1[4 * r + c] = projMatrl.at<double>(r, c);
            p2[4 * r + c] = projMatrr.at<double>(r, c);
        }
    svo_ctx* c = mpORBextractorLeft_->context(current_frame_->left_img_.cols, current_frame_->left_img_.rows);
    if (svo_triangulate_points(c, p1, p2, &points_t1_left[0].x, &points_t1_right[0].x, n, nullptr, points3D_t1.ptr<float>()) != SVO_OK)
        throw std::runtime_error(std::string("svo_triangulate_points: ") + svo_last_error(c));
}

bool Tracking::OpenCV_EstimatePose_PnP(cv::Mat& projMatrl, std::vector<cv::Point2f>& pointsLeft_t2, cv::Mat& points3D_t1, cv::Mat& rotation,
                                       cv::Mat& translation) {  // tracking.cpp:464-501
    const int n = (int)pointsLeft_t2.size();
    rotation.create(3, 3, CV_64F);
    translation.create(3, 1, CV_64F);
    last_inliers_.clear();
    double K[9], rvec[3] = {0, 0, 0}, tvec[3] = {0, 0, 0}, R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) K[3 * r + c] = projMatrl.at<double>(r, c);  // intrinsic_matrix (:476-477)
    int32_t ok = 0, n_inliers = 0;
    if (n > 0) {
        // points3D_t1: n x 1 CV_32FC3 (or n x 3 CV_32F), gathered row by row so that a non-continuous Mat works too
        std::vector<float> p3((size_t)3 * n);
        for (int i = 0; i < n; ++i) {
            const float* row = points3D_t1.ptr<float>(i);
            p3[3 * i] = row[0];
            p3[3 * i + 1] = row[1];
            p3[3 * i + 2] = row[2];
        }
        std::vector<uint8_t> mask((size_t)n);
        svo_ctx* c = mpORBextractorLeft_->context(current_frame_->left_img_.cols, current_frame_->left_img_.rows);
        const int32_t count = n;
        if (svo_pnp_ransac(c, K, p3.data(), &pointsLeft_t2[0].x, &count, 1, n, iterationsCount_, reprojectionError_, (double)confidence_, rvec, tvec, R,
                           mask.data(), &n_inliers, nullptr, &ok) != SVO_OK)
            throw std::runtime_error(std::string("svo_pnp_ransac: ") + svo_last_error(c));
        for (int i = 0; i < n; ++i)
            if (mask[i]) last_inliers_.push_back(i);
    }
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) rotation.at<double>(r, c) = R[3 * r + c];  // cv::Rodrigues(rvec, rotation) (:489)
        translation.at<double>(r, 0) = tvec[r];
    }
    if (n == 0) return false;
    return !((double)n_inliers / (double)n < inlier_rate_);  // :492-500
}

int Tracking::Detect_OpenCVFASTFeatures() {  // tracking.cpp:94-113
    const cv::Mat& img = current_frame_->left_img_;
    if (img.empty()) return 0;
    svo_ctx* c = mpORBextractorLeft_->context(img.cols, img.rows);
    const int cap = std::max(4096, img.cols * img.rows / 16);
    std::vector<cv::KeyPoint> keypoints((size_t)cap);
    int n = 0;
    const int fast_threshold = 20;
    if (svo_fast_detect(c, img.data, img.cols, img.rows, (size_t)img.step, fast_threshold, 1, reinterpret_cast<svo_keypoint*>(keypoints.data()), cap, &n) != SVO_OK)
        throw std::runtime_error(std::string("svo_fast_detect: ") + svo_last_error(c));
    Feature::wrap_keypoints(current_frame_, keypoints.data(), (size_t)n, true, current_frame_->features_left_);  // one block, not n `new`s
    return n;
}

int Tracking::Detect_MyORBFeatures() {  // tracking.cpp:502-532
    std::vector<cv::KeyPoint> mvKeys1, mvKeys2;
    cv::Mat mDescriptors1, mDescriptors2;
    (*mpORBextractorLeft_)(current_frame_->left_img_, cv::Mat(), mvKeys1, mDescriptors1);
    (*mpORBextractorLeft_)(current_frame_->right_img_, cv::Mat(), mvKeys2, mDescriptors2);
    current_frame_->left_Descriptors_ = mDescriptors1;
    current_frame_->right_Descriptors_ = mDescriptors2;
    // the reference sets is_on_left_image_ for neither side (default true); kept
    Feature::wrap_keypoints(current_frame_, mvKeys1.data(), mvKeys1.size(), true, current_frame_->features_left_);
    Feature::wrap_keypoints(current_frame_, mvKeys2.data(), mvKeys2.size(), true, current_frame_->features_right_);
    return (int)std::min(mvKeys1.size(), mvKeys2.size());
}

static void bf_match(svo_ctx* c, const cv::Mat& q, const cv::Mat& t, std::vector<cv::DMatch>& out) {
    out.clear();
    if (q.empty() || t.empty()) return;  // OpenCV returns no matches for an empty train set
    std::vector<int32_t> idx((size_t)q.rows), dist((size_t)q.rows);
    if (!q.isContinuous() || !t.isContinuous()) throw std::runtime_error("descriptor matrices must be continuous");
    if (svo_match_hamming(c, q.data, q.rows, t.data, t.rows, idx.data(), dist.data()) != SVO_OK)
        throw std::runtime_error(std::string("svo_match_hamming: ") + svo_last_error(c));
    out.reserve((size_t)q.rows);
    for (int i = 0; i < q.rows; ++i) out.push_back(cv::DMatch(i, idx[i], (float)dist[i]));
}

int Tracking::ORB_Robust_Find_MuliImage_MatchedFeatures(std::vector<cv::Point2f>& points_t2_left, std::vector<cv::Point2f>& points_t1_left,
                                                        std::vector<cv::Point2f>& points_t1_right) {  // tracking.cpp:534-581
    svo_ctx* c = mpORBextractorLeft_->context(current_frame_->left_img_.cols, current_frame_->left_img_.rows);
    std::vector<cv::DMatch>& match1 = last_match1_;
    std::vector<cv::DMatch>& match2 = last_match2_;
    bf_match(c, last_frame_->left_Descriptors_, last_frame_->right_Descriptors_, match1);
    bf_match(c, last_frame_->left_Descriptors_, current_frame_->left_Descriptors_, match2);
    // the association filter (tracking.cpp:546-577) runs on the device: the same kernel (K10) as in the throughput mode
    auto kps_of = [](const std::vector<Feature::Ptr>& f) {
        std::vector<cv::KeyPoint> k;
        k.reserve(f.size());
        for (const auto& p : f) k.push_back(p->position_);
        return k;
    };
    const std::vector<cv::KeyPoint> kll = kps_of(last_frame_->features_left_), klr = kps_of(last_frame_->features_right_),
                                    kcl = kps_of(current_frame_->features_left_);
    const int nq = std::min((int)match1.size(), (int)match2.size());
    std::vector<int32_t> i1((size_t)nq), d1((size_t)nq), i2((size_t)nq), d2((size_t)nq);
    for (int i = 0; i < nq; ++i) {
        i1[i] = match1[i].trainIdx;
        d1[i] = (int32_t)match1[i].distance;
        i2[i] = match2[i].trainIdx;
        d2[i] = (int32_t)match2[i].distance;
    }
    // queries beyond the matches' length cannot be filtered (an empty train set yields no matches): same early-out as the batch path
    const int n_ll = std::min((int)kll.size(), nq);
    const int cap = std::max(1, n_ll);
    std::vector<cv::Point2f> t2((size_t)cap), t1l((size_t)cap), t1r((size_t)cap);
    int n = 0;
    if (svo_orb_match_filter(c, reinterpret_cast<const svo_keypoint*>(kll.data()), n_ll, reinterpret_cast<const svo_keypoint*>(klr.data()), (int)klr.size(),
                             reinterpret_cast<const svo_keypoint*>(kcl.data()), (int)kcl.size(), i1.data(), d1.data(), i2.data(), d2.data(),
                             feature_match_error_, reinterpret_cast<float*>(t2.data()), reinterpret_cast<float*>(t1l.data()),
                             reinterpret_cast<float*>(t1r.data()), cap, &n) != SVO_OK)
        throw std::runtime_error(std::string("svo_orb_match_filter: ") + svo_last_error(c));
    points_t2_left.insert(points_t2_left.end(), t2.begin(), t2.begin() + n);
    points_t1_left.insert(points_t1_left.end(), t1l.begin(), t1l.begin() + n);
    points_t1_right.insert(points_t1_right.end(), t1r.begin(), t1r.begin() + n);
    return (int)points_t2_left.size();
}

static void lk_call(svo_ctx* c, const cv::Mat& prev, const cv::Mat& next, const std::vector<cv::Point2f>& p, std::vector<cv::Point2f>& q,
                    std::vector<unsigned char>& status) {
    q.assign(p.size(), cv::Point2f());
    status.assign(p.size(), 0);
    if (p.empty()) return;
    std::vector<float> err(p.size());
    if ((size_t)prev.step != (size_t)next.step) throw std::runtime_error("stereo images must share a row stride");
    if (svo_lk_track(c, prev.data, next.data, prev.cols, prev.rows, (size_t)prev.step, reinterpret_cast<const float*>(p.data()), (int)p.size(),
                     reinterpret_cast<float*>(q.data()), status.data(), err.data(), 21, 3, 30, 0.01, 0.001) != SVO_OK)
        throw std::runtime_error(std::string("svo_lk_track: ") + svo_last_error(c));
}

int Tracking::LK_Robust_Find_MuliImage_MatchedFeatures(std::vector<cv::Point2f>& points_t2_left, std::vector<cv::Point2f>& points_t2_right,
                                                       std::vector<cv::Point2f>& points_t1_left, std::vector<cv::Point2f>& points_t1_right) {
    // tracking.cpp:583-622: four chained calls, then the joint filter
    svo_ctx* c = mpORBextractorLeft_->context(current_frame_->left_img_.cols, current_frame_->left_img_.rows);
    std::vector<cv::Point2f> points_final;
    std::vector<unsigned char> status1, status2, status3, status4;
    lk_call(c, last_frame_->left_img_, last_frame_->right_img_, points_t1_left, points_t1_right, status1);
    lk_call(c, last_frame_->right_img_, current_frame_->right_img_, points_t1_right, points_t2_right, status2);
    lk_call(c, current_frame_->right_img_, current_frame_->left_img_, points_t2_right, points_t2_left, status3);
    lk_call(c, current_frame_->left_img_, last_frame_->left_img_, points_t2_left, points_final, status4);
    deleteBadmatchFeatures(points_t1_left, points_t1_right, points_t2_right, points_t2_left, points_final, status1, status2, status3, status4);
    return (int)points_t1_left.size();
}

void Tracking::deleteBadmatchFeatures(std::vector<cv::Point2f>& points0, std::vector<cv::Point2f>& points1, std::vector<cv::Point2f>& points2,
                                      std::vector<cv::Point2f>& points3, std::vector<cv::Point2f>& points0_return,
                                      std::vector<unsigned char>& status0, std::vector<unsigned char>& status1,
                                      std::vector<unsigned char>& status2, std::vector<unsigned char>& status3) {
    // tracking.cpp:623-660 on the device: the same kernel (K11) as in the throughput mode; the five lists are compacted in place
    const size_t n = status3.size();
    if (n == 0) return;
    if (points0.size() < n || points1.size() < n || points2.size() < n || points3.size() < n || points0_return.size() < n || status0.size() < n ||
        status1.size() < n || status2.size() < n)
        throw std::runtime_error("deleteBadmatchFeatures: lists shorter than status3");
    if (!current_frame_) throw std::runtime_error("deleteBadmatchFeatures: no current frame");
    svo_ctx* c = mpORBextractorLeft_->context(current_frame_->left_img_.cols, current_frame_->left_img_.rows);
    int kept = 0;
    if (svo_lk_chain_filter(c, (int)n, reinterpret_cast<float*>(points0.data()), reinterpret_cast<float*>(points1.data()),
                            reinterpret_cast<float*>(points2.data()), reinterpret_cast<float*>(points3.data()),
                            reinterpret_cast<float*>(points0_return.data()), status0.data(), status1.data(), status2.data(), status3.data(),
                            feature_match_error_, &kept) != SVO_OK)
        throw std::runtime_error(std::string("svo_lk_chain_filter: ") + svo_last_error(c));
    points0.resize((size_t)kept);
    points1.resize((size_t)kept);
    points2.resize((size_t)kept);
    points3.resize((size_t)kept);
    points0_return.resize((size_t)kept);
}

}  // namespace lzb_vio
