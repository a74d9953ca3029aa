// include/lzb_vio/feature.h -- boundary type of the drop-in build.
//
// The reference wraps every detected keypoint in its own heap object: `Feature::Ptr(new Feature(frame, kp))`, ~2 k allocations
// per stereo frame (tracking.cpp:105-110, 519-529; SURVEY.md 8a row a13 calls it pure overhead).  Callers of the reference's API
// read `features_left_[i]->position_` (tracking.cpp:566-571), so the element type has to stay a std::shared_ptr<Feature> with
// these field names -- but nothing requires one allocation per keypoint.  wrap_keypoints() below builds all features of an
// image in ONE block and hands out aliasing shared_ptrs into it; the block lives as long as any of its features does.
#pragma once
#include <cstddef>
#include <memory>
#include <vector>

#ifdef SVO_USE_REAL_OPENCV
#include <opencv2/features2d.hpp>
#else
#include "cvcompat/opencv2/core.hpp"
#endif

namespace lzb_vio {

struct Frame;

struct Feature {
    typedef std::shared_ptr<Feature> Ptr;

    // same public data members as the reference's struct (feature.h:22-26), same defaults
    std::weak_ptr<Frame> frame_;
    cv::KeyPoint position_;
    cv::Mat Descriptor_;  // never filled by the reference either (descriptors live in Frame::left_/right_Descriptors_)
    bool is_outlier_ = false;
    bool is_on_left_image_ = true;

    Feature() = default;
    Feature(std::shared_ptr<Frame> frame, const cv::KeyPoint& kp) : frame_(std::move(frame)), position_(kp) {}

    // Appends one Feature per keypoint of `kps[0 .. n)` to `out`, all stored in a single shared block.
    static void wrap_keypoints(const std::shared_ptr<Frame>& frame, const cv::KeyPoint* kps, std::size_t n, bool on_left,
                               std::vector<Ptr>& out) {
        if (n == 0) return;
        auto block = std::make_shared<std::vector<Feature>>(n);
        out.reserve(out.size() + n);
        for (std::size_t i = 0; i < n; ++i) {
            Feature& f = (*block)[i];
            f.frame_ = frame;
            f.position_ = kps[i];
            f.is_on_left_image_ = on_left;
            out.push_back(Ptr(block, &f));  // aliasing constructor: shares ownership of the block, points at element i
        }
    }
};

}  // namespace lzb_vio
