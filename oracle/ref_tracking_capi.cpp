// oracle/ref_tracking_capi.cpp -- TEST INFRASTRUCTURE ONLY.  Hosts the reference's own
//     Tracking::ORB_Robust_Find_MuliImage_MatchedFeatures, Tracking::LK_Robust_Find_MuliImage_MatchedFeatures and
//     Tracking::deleteBadmatchFeatures                                   (/root/reference/src/tracking.cpp:534-660)
// as compiled code.  tracking.cpp as a whole needs Eigen, Sophus, g2o, glog and the Frame / Map graph, none of which exist
// in this image, but these three member functions only touch a handful of members.  The recipe (oracle/Makefile, target
// `ref`) cuts their text out of the reference file at BUILD time into oracle/_ref/tracking_slice.inc (git-ignored, never
// committed) and this translation unit supplies the minimal declarations they need -- the same member names and types as
// include/lzb_vio/{frame,feature,tracking}.h (frame.h:32-37, feature.h:23, tracking.h:52-59,91-95,108-109,136) -- with the
// OpenCV calls served by oracle/ref_shim (BF-Hamming match and calcOpticalFlowPyrLK -> cv2-pinned restatements).
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>

#include "shim_cv.hpp"

// The slice prints two or three lines per call (tracking.cpp:560-561, 579: printf and std::cout).  They are routed to a sink so
// that test logs stay readable and the library needs no iostream state (this toolchain links libstdc++ statically, and a
// second copy of std::cout inside a dlopen'ed library is not safe in a process that already holds the dynamic one).
namespace std {
struct RefNullStream {
    template <typename T>
    RefNullStream& operator<<(const T&) { return *this; }
};
static RefNullStream ref_null_stream;
static const int ref_endl = 0;
}  // namespace std
#define cout ref_null_stream
#define endl ref_endl
#define printf(...) ((void)0)

namespace lzb_vio {

struct Feature {
    typedef std::shared_ptr<Feature> Ptr;
    cv::KeyPoint position_;
};

struct Frame {
    typedef std::shared_ptr<Frame> Ptr;
    cv::Mat left_img_, right_img_;
    cv::Mat left_Descriptors_, right_Descriptors_;
    std::vector<std::shared_ptr<Feature>> features_left_, features_right_;
};

class Tracking {
public:
    int LK_Robust_Find_MuliImage_MatchedFeatures(std::vector<cv::Point2f>& points_t2_left, std::vector<cv::Point2f>& points_t2_right,
                                                 std::vector<cv::Point2f>& points_t1_left, std::vector<cv::Point2f>& points_t1_right);
    int ORB_Robust_Find_MuliImage_MatchedFeatures(std::vector<cv::Point2f>& points_t2_left, std::vector<cv::Point2f>& points_t1_left,
                                                  std::vector<cv::Point2f>& points_t1_right);
    void deleteBadmatchFeatures(std::vector<cv::Point2f>& points0, std::vector<cv::Point2f>& points1, std::vector<cv::Point2f>& points2,
                                std::vector<cv::Point2f>& points3, std::vector<cv::Point2f>& points0_return, std::vector<cv::uchar>& status0,
                                std::vector<cv::uchar>& status1, std::vector<cv::uchar>& status2, std::vector<cv::uchar>& status3);
    Frame::Ptr current_frame_ = nullptr;
    Frame::Ptr last_frame_ = nullptr;
    double feature_match_error_ = 10;
};

using cv::uchar;
#include "_ref/tracking_slice.inc"  // the three member function definitions, verbatim from the reference at build time

}  // namespace lzb_vio

namespace {

lzb_vio::Frame::Ptr make_frame(const uint8_t* left, const uint8_t* right, int w, int h, const void* kl, const uint8_t* dl, int nl, const void* kr,
                               const uint8_t* dr, int nr) {
    auto f = std::make_shared<lzb_vio::Frame>();
    if (left) f->left_img_ = cv::Mat(h, w, cv::CV_8UC1, const_cast<uint8_t*>(left), (size_t)w);
    if (right) f->right_img_ = cv::Mat(h, w, cv::CV_8UC1, const_cast<uint8_t*>(right), (size_t)w);
    auto fill = [](const void* k, const uint8_t* d, int n, std::vector<std::shared_ptr<lzb_vio::Feature>>& feats, cv::Mat& desc) {
        const cv::KeyPoint* kp = static_cast<const cv::KeyPoint*>(k);
        for (int i = 0; i < n; ++i) {
            auto ft = std::make_shared<lzb_vio::Feature>();
            ft->position_ = kp[i];
            feats.push_back(ft);
        }
        if (n) desc = cv::Mat(n, 32, cv::CV_8UC1, const_cast<uint8_t*>(d), 32);
    };
    fill(kl, dl, nl, f->features_left_, f->left_Descriptors_);
    fill(kr, dr, nr, f->features_right_, f->right_Descriptors_);
    return f;
}

int copy_out(const std::vector<cv::Point2f>& v, float* dst, int cap) {
    const int n = (int)v.size();
    if (n <= cap && n) std::memcpy(dst, v.data(), sizeof(cv::Point2f) * (size_t)n);
    return n;
}

}  // namespace

extern "C" {

// Tracking::ORB_Robust_Find_MuliImage_MatchedFeatures on (last frame: left/right keypoints + descriptors; current frame: left
// keypoints + descriptors).  Returns the number of accepted matches; the three point lists are cap x 2 floats.
int ref_orb_robust_match(const void* last_kl, const uint8_t* last_dl, int n_ll, const void* last_kr, const uint8_t* last_dr, int n_lr,
                         const void* cur_kl, const uint8_t* cur_dl, int n_cl, double feature_match_error, float* t2_left, float* t1_left,
                         float* t1_right, int cap) {
    lzb_vio::Tracking t;
    t.feature_match_error_ = feature_match_error;
    t.last_frame_ = make_frame(nullptr, nullptr, 0, 0, last_kl, last_dl, n_ll, last_kr, last_dr, n_lr);
    t.current_frame_ = make_frame(nullptr, nullptr, 0, 0, cur_kl, cur_dl, n_cl, nullptr, nullptr, 0);
    std::vector<cv::Point2f> a, b, c;
    t.ORB_Robust_Find_MuliImage_MatchedFeatures(a, b, c);
    copy_out(a, t2_left, cap);
    copy_out(b, t1_left, cap);
    return copy_out(c, t1_right, cap);
}

// Tracking::LK_Robust_Find_MuliImage_MatchedFeatures: the four chained calcOpticalFlowPyrLK calls + deleteBadmatchFeatures.
// seeds: n x 2 floats (points_t1_left on entry).  Returns the number of surviving tracks; outputs are cap x 2 floats.
int ref_lk_robust_track(const uint8_t* last_left, const uint8_t* last_right, const uint8_t* cur_left, const uint8_t* cur_right, int w, int h,
                        const float* seeds, int n, double feature_match_error, float* t2_left, float* t2_right, float* t1_left, float* t1_right,
                        int cap) {
    lzb_vio::Tracking t;
    t.feature_match_error_ = feature_match_error;
    t.last_frame_ = make_frame(last_left, last_right, w, h, nullptr, nullptr, 0, nullptr, nullptr, 0);
    t.current_frame_ = make_frame(cur_left, cur_right, w, h, nullptr, nullptr, 0, nullptr, nullptr, 0);
    std::vector<cv::Point2f> p_t2l, p_t2r, p_t1l((size_t)n), p_t1r;
    if (n) std::memcpy(p_t1l.data(), seeds, sizeof(cv::Point2f) * (size_t)n);
    t.LK_Robust_Find_MuliImage_MatchedFeatures(p_t2l, p_t2r, p_t1l, p_t1r);
    copy_out(p_t2l, t2_left, cap);
    copy_out(p_t2r, t2_right, cap);
    copy_out(p_t1l, t1_left, cap);
    return copy_out(p_t1r, t1_right, cap);
}
}
