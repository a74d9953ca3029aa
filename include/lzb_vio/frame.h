// include/lzb_vio/frame.h -- boundary type, same public fields the front-end touches as the reference's
// include/lzb_vio/frame.h:20-65 (left_img_/right_img_, left_/right_Descriptors_, features_left_/right_).  The pose
// member is Sophus::SE3d in the reference; Sophus is not part of the hot path and not in this image, so the pose is kept
// as a row-major 4x4 double matrix unless SVO_HAVE_SOPHUS is defined.
#pragma once
#include <array>
#include <memory>
#include <mutex>
#include <vector>

#ifdef SVO_USE_REAL_OPENCV
#include <opencv2/core/core.hpp>
#else
#include "cvcompat/opencv2/core.hpp"
#endif
#ifdef SVO_HAVE_SOPHUS
#include <sophus/se3.hpp>
#endif

namespace lzb_vio {

struct Feature;

#ifdef SVO_HAVE_SOPHUS
typedef Sophus::SE3d SE3;
#else
struct SE3 {
    std::array<double, 16> m{{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}};
};
#endif

struct Frame {
    typedef std::shared_ptr<Frame> Ptr;

    unsigned long id_ = 0;
    unsigned long keyframe_id_ = 0;
    bool is_keyframe_ = false;
    double time_stamp_ = 0;
    SE3 pose_;
    std::mutex pose_mutex_;
    cv::Mat left_img_, right_img_;
    cv::Mat left_Descriptors_, right_Descriptors_;
    std::vector<std::shared_ptr<Feature>> features_left_;
    std::vector<std::shared_ptr<Feature>> features_right_;
    std::vector<unsigned char> status_;

    Frame() {}
    Frame(long id, double time_stamp, const SE3& pose, const cv::Mat& left, const cv::Mat& right)
        : id_((unsigned long)id), time_stamp_(time_stamp), pose_(pose), left_img_(left), right_img_(right) {}

    SE3 Pose() {
        std::unique_lock<std::mutex> lck(pose_mutex_);
        return pose_;
    }
    void SetPose(const SE3& pose) {
        std::unique_lock<std::mutex> lck(pose_mutex_);
        pose_ = pose;
    }
    static std::shared_ptr<Frame> CreateFrame() {
        static unsigned long factory_id = 0;
        Frame::Ptr f(new Frame);
        f->id_ = factory_id++;
        return f;
    }
};

}  // namespace lzb_vio
