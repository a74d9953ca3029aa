// include/lzb_vio/feature.h -- boundary type, same public fields as the reference's include/lzb_vio/feature.h:16-33.
#pragma once
#include <memory>

#ifdef SVO_USE_REAL_OPENCV
#include <opencv2/features2d.hpp>
#else
#include "cvcompat/opencv2/core.hpp"
#endif

namespace lzb_vio {

struct Frame;

struct Feature {
    typedef std::shared_ptr<Feature> Ptr;
    std::weak_ptr<Frame> frame_;    // frame that holds this feature
    cv::KeyPoint position_;         // 2-D detection
    cv::Mat Descriptor_;            // unused by the reference (quirk Q6)
    bool is_outlier_ = false;
    bool is_on_left_image_ = true;

    Feature() {}
    Feature(std::shared_ptr<Frame> frame, const cv::KeyPoint& kp) : frame_(frame), position_(kp) {}
};

}  // namespace lzb_vio
