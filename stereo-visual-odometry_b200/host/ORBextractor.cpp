// host/ORBextractor.cpp -- lzb_vio::ORBextractor as a thin host wrapper over the C ABI (include/svo_cuda.h).
// Replaces the reference's src/ORBextractor.cpp; no pixel arithmetic happens on the host.
#include "lzb_vio/ORBextractor.h"

#include <algorithm>
#include <cassert>
#include <cstring>
#include <stdexcept>

#include "svo_cuda.h"

namespace lzb_vio {

int ORBextractor::default_device_ = 0;

static const int kEdge = 19;  // EDGE_THRESHOLD (ORBextractor.cpp:19)

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST), minThFAST(_minThFAST) {
    device_ = default_device_;
    // the tables come from the library so that host getters and device kernels can never disagree
    svo_ctx* probe = nullptr;
    if (svo_ctx_create(&probe, device_, 64, 64, 2) != SVO_OK) throw std::runtime_error("lzb_vio::ORBextractor: no CUDA device (there is no CPU fallback)");
    if (svo_orb_configure(probe, nfeatures, _scaleFactor, nlevels, iniThFAST, minThFAST) != SVO_OK) {
        std::string msg = svo_last_error(probe);
        svo_ctx_destroy(probe);
        throw std::runtime_error("lzb_vio::ORBextractor: " + msg);
    }
    mvScaleFactor.resize(nlevels);
    mvInvScaleFactor.resize(nlevels);
    mvLevelSigma2.resize(nlevels);
    mvInvLevelSigma2.resize(nlevels);
    mnFeaturesPerLevel.resize(nlevels);
    svo_orb_tables(probe, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(), mvInvLevelSigma2.data(), mnFeaturesPerLevel.data());
    svo_ctx_destroy(probe);
    mvImagePyramid.resize(nlevels);
}

ORBextractor::~ORBextractor() {
    if (ctx_) svo_ctx_destroy(ctx_);
}

svo_ctx* ORBextractor::context(int w, int h) {
    if (ctx_ && w <= ctx_w_ && h <= ctx_h_) return ctx_;
    if (ctx_) svo_ctx_destroy(ctx_);
    ctx_ = nullptr;
    ctx_w_ = std::max(w, ctx_w_);
    ctx_h_ = std::max(h, ctx_h_);
    if (svo_ctx_create(&ctx_, device_, ctx_w_, ctx_h_, 2) != SVO_OK) throw std::runtime_error("lzb_vio::ORBextractor: cannot create the CUDA context");
    if (svo_orb_configure(ctx_, nfeatures, (float)scaleFactor, nlevels, iniThFAST, minThFAST) != SVO_OK)
        throw std::runtime_error(std::string("lzb_vio::ORBextractor: ") + svo_last_error(ctx_));
    return ctx_;
}

// mvImagePyramid[l] = level l inside a 19-px BORDER_REFLECT_101 frame (ComputePyramid, ORBextractor.cpp:1061-1085).  The
// border is a pure copy of level pixels, made on the host from the level read back from the device.
void ORBextractor::fill_pyramid(int w, int h) {
    for (int l = 0; l < nlevels; ++l) {
        int lw = 0, lh = 0;
        if (svo_orb_level_size(ctx_, w, h, l, &lw, &lh) != SVO_OK) return;
        cv::Mat whole(lh + 2 * kEdge, lw + 2 * kEdge, CV_8UC1);
        unsigned char* inner = whole.ptr(kEdge) + kEdge;
        if (svo_orb_read_level(ctx_, 0, l, inner, (size_t)whole.step, &lw, &lh) != SVO_OK) return;
        auto refl = [](int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); };
        for (int y = 0; y < lh; ++y) {
            unsigned char* row = inner + (size_t)y * (size_t)whole.step;
            for (int x = 1; x <= kEdge; ++x) {
                row[-x] = row[refl(-x, lw)];
                row[lw - 1 + x] = row[refl(lw - 1 + x, lw)];
            }
        }
        for (int y = 1; y <= kEdge; ++y) {
            std::memcpy(whole.ptr(kEdge - y), whole.ptr(kEdge + refl(-y, lh)), (size_t)whole.step);
            std::memcpy(whole.ptr(kEdge + lh - 1 + y), whole.ptr(kEdge + refl(lh - 1 + y, lh)), (size_t)whole.step);
        }
        // the reference stores the ROI without the border (temp(Rect(19,19,w,h)), :1069); keep that view semantics
        mvImagePyramid[l] = cv::Mat(lh, lw, CV_8UC1, inner, (size_t)whole.step);
        pyramid_store_[l] = whole;
    }
}

void ORBextractor::operator()(cv::InputArray _image, cv::InputArray /*mask*/, std::vector<cv::KeyPoint>& keypoints, cv::OutputArray _descriptors) {
    if (_image.empty()) return;
    const cv::Mat image = _image.getMat();  // ORBextractor.cpp:996
    assert(image.type() == CV_8UC1);
    const int w = image.cols, h = image.rows;
    svo_ctx* c = context(w, h);
    const int cap = svo_orb_capacity(c, w, h);
    kp_buf_.resize((size_t)std::max(cap, 1));
    desc_buf_.resize((size_t)std::max(cap, 1) * 32);
    int n = 0;
    static_assert(sizeof(cv::KeyPoint) == sizeof(svo_keypoint), "cv::KeyPoint layout");
    const int rc = svo_orb_extract(c, image.data, w, h, (size_t)image.step, reinterpret_cast<svo_keypoint*>(kp_buf_.data()), desc_buf_.data(), cap, &n);
    if (rc != SVO_OK) throw std::runtime_error(std::string("svo_orb_extract: ") + svo_last_error(c));
    if (n == 0) {
        _descriptors.release();
    } else {
        _descriptors.create(n, 32, CV_8U);  // ORBextractor.cpp:1017
        cv::Mat descriptors = _descriptors.getMat();
        for (int i = 0; i < n; ++i) std::memcpy(descriptors.ptr(i), desc_buf_.data() + (size_t)i * 32, 32);
    }
    keypoints.assign(kp_buf_.begin(), kp_buf_.begin() + n);
    if (export_pyramid_) {
        pyramid_store_.resize(nlevels);
        fill_pyramid(w, h);
    }
}

}  // namespace lzb_vio
