// host/tracking_frontend.cpp -- the front-end half of lzb_vio::Tracking over the C ABI.
// Replaces the OpenCV calls at tracking.cpp:101 (cv::FAST), :508-509 (ORBextractor), :539-544 (BFMatcher), :593-618
// (calcOpticalFlowPyrLK); the filters around them (:546-577, :623-660) are restated on the host as in the reference.
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
}

Tracking::~Tracking() { delete mpORBextractorLeft_; }

void Tracking::Readparameter() {  // tracking.cpp:22-47, front-end fields
    feature_match_error_ = parameter_->feature_match_error_;
    track_mode_ = parameter_->track_mode_;
    num_features_tracking_ = parameter_->num_features_tracking_;
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
    return pose_backend_ ? pose_backend_(matched_t1_left, matched_t1_right, matched_t2_left) : true;
}

bool Tracking::LK_StereoF2F_PnP_Track() {  // tracking.cpp:258-288
    const int num_features_left = Detect_OpenCVFASTFeatures();
    if (num_features_left < 30) return false;
    std::vector<cv::Point2f> matched_t2_left, matched_t1_left, matched_t2_right, matched_t1_right;
    for (auto& feature : last_frame_->features_left_) matched_t1_left.push_back(feature->position_.pt);
    const int n = LK_Robust_Find_MuliImage_MatchedFeatures(matched_t2_left, matched_t2_right, matched_t1_left, matched_t1_right);
    if (n < num_features_tracking_) return false;
    return pose_backend_ ? pose_backend_(matched_t1_left, matched_t1_right, matched_t2_left) : true;
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
    const int des_index =
        std::min(std::min(last_frame_->left_Descriptors_.rows, last_frame_->right_Descriptors_.rows), current_frame_->left_Descriptors_.rows);
    if ((int)match1.size() < des_index || (int)match2.size() < des_index) return 0;
    double min_dist = 10000, max_dist = 0;
    for (int i = 0; i < des_index; i++) {
        double dist = std::max(match1[i].distance, match2[i].distance);
        if (dist < min_dist) min_dist = dist;
        if (dist > max_dist) max_dist = dist;
    }
    for (int i = 0; i < des_index; i++) {
        if ((match1[i].distance <= std::max(2 * min_dist, 30.0)) && (match2[i].distance <= std::max(2 * min_dist, 30.0)) &&
            (std::abs((last_frame_->features_left_[match1[i].queryIdx]->position_.pt.y) -
                      (last_frame_->features_right_[match1[i].trainIdx]->position_.pt.y)) < feature_match_error_)) {
            points_t2_left.push_back((current_frame_->features_left_)[match2[i].trainIdx]->position_.pt);
            points_t1_left.push_back((last_frame_->features_left_)[match1[i].queryIdx]->position_.pt);
            points_t1_right.push_back((last_frame_->features_right_)[match1[i].trainIdx]->position_.pt);
        }
    }
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
    // tracking.cpp:623-660.  Same keep/drop rule; the reference erases in place (O(n^2)), here survivors are compacted.
    size_t w = 0;
    for (size_t i = 0; i < status3.size(); i++) {
        const cv::Point2f pt0 = points0[i], pt1 = points1[i], pt2 = points2[i], pt3 = points3[i], pt0_r = points0_return[i];
        const bool outsideframe = (pt3.x < 0) || (pt3.y < 0) || (pt2.x < 0) || (pt2.y < 0) || (pt1.x < 0) || (pt1.y < 0) || (pt0.x < 0) ||
                                  (pt0.y < 0) || (pt0_r.x < 0) || (pt0_r.y < 0);
        const bool badstatus = (status2[i] == 0) || (status1[i] == 0) || (status0[i] == 0) || (status3[i] == 0);
        const bool NoStereoEpipolarConstrain =
            (std::abs(pt0.y - pt1.y) > feature_match_error_) || (std::abs(pt2.y - pt3.y) > feature_match_error_);
        if (outsideframe || badstatus || NoStereoEpipolarConstrain) continue;
        points0[w] = pt0;
        points1[w] = pt1;
        points2[w] = pt2;
        points3[w] = pt3;
        points0_return[w] = pt0_r;
        ++w;
    }
    points0.resize(w);
    points1.resize(w);
    points2.resize(w);
    points3.resize(w);
    points0_return.resize(w);
}

}  // namespace lzb_vio
