// include/lzb_vio/ORBextractor.h -- drop-in for the reference's include/lzb_vio/ORBextractor.h:24-100.
//
// Same class name, constructor, operator() and getters; the body calls libsvo_b200.so (include/svo_cuda.h) instead of
// OpenCV: pyramid, per-cell FAST, octree, IC_Angle and steered BRIEF all run as sm_100a kernels.  There is no CPU
// fallback: if no CUDA device is present the constructor throws std::runtime_error.
#pragma once
#include <string>
#include <vector>

#ifdef SVO_USE_REAL_OPENCV
#include <opencv2/core/core.hpp>
#include <opencv2/features2d/features2d.hpp>
#else
#include "cvcompat/opencv2/core.hpp"
#endif

struct svo_ctx;

namespace lzb_vio {

class ORBextractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    // reference: ORBextractor.h:33-34
    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
    ~ORBextractor();
    ORBextractor(const ORBextractor&) = delete;
    ORBextractor& operator=(const ORBextractor&) = delete;

    // reference: ORBextractor.h:41-43.  The mask is ignored, as in the reference.  An empty image returns with the
    // outputs untouched; a non-8UC1 image fails the same live assert (ORBextractor.cpp:993-997).
    void operator()(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint>& keypoints, cv::OutputArray descriptors);

    int GetLevels() { return nlevels; }
    float GetScaleFactor() { return (float)scaleFactor; }
    std::vector<float> GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

    // Public member of the reference (ORBextractor.h:75): the pyramid of the last image, each level inside a 19-px
    // BORDER_REFLECT_101 frame.  Filled after every operator() unless SetExportPyramid(false) (throughput setting: saves a
    // 1.4 MB device-to-host copy per 1241x376 image; nothing in the reference reads this member).
    std::vector<cv::Mat> mvImagePyramid;
    void SetExportPyramid(bool on) { export_pyramid_ = on; }

    // ---- additions for the Tracking front-end (not in the reference)
    svo_ctx* context(int w, int h);  // the CUDA context, created / grown on demand for w x h images
    int device() const { return device_; }
    static void SetDevice(int device) { default_device_ = device; }

protected:
    int nfeatures;
    double scaleFactor;
    int nlevels;
    int iniThFAST;
    int minThFAST;
    std::vector<int> mnFeaturesPerLevel;
    std::vector<float> mvScaleFactor;
    std::vector<float> mvInvScaleFactor;
    std::vector<float> mvLevelSigma2;
    std::vector<float> mvInvLevelSigma2;

private:
    void fill_pyramid(int w, int h);
    svo_ctx* ctx_ = nullptr;
    int ctx_w_ = 0, ctx_h_ = 0, device_ = 0;
    bool export_pyramid_ = true;
    std::vector<cv::KeyPoint> kp_buf_;
    std::vector<unsigned char> desc_buf_;
    std::vector<cv::Mat> pyramid_store_;  // owners of the bordered levels mvImagePyramid views into
    static int default_device_;
};

}  // namespace lzb_vio
