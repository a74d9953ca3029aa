// oracle/ref_shim/shim_cv.hpp -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
//
// A header stand-in for the slice of the OpenCV C++ API that /root/reference/src/ORBextractor.cpp uses, so that the
// reference's OWN source file can be compiled here, unmodified, where it lies (recipe: oracle/Makefile, target _ref).
// This image has no OpenCV C++ headers or libraries; the containers (Mat with shared ROI views, KeyPoint,
// Point_, Size, Rect, InputArray / OutputArray proxies) are written out below, and the five OpenCV ALGORITHMS on the path
// (resize INTER_LINEAR, copyMakeBorder, FAST, GaussianBlur 7x7 sigma 2, fastAtan2) forward to the oracle's restatements
// in oracle_prims.cpp, each of which is pinned bit-for-bit against cv2 4.13.0 (tests/test_oracle_cpu.py, tests/golden/).
// What this buys: the reference's own loops -- cell grid, two-threshold FAST fallback, DistributeOctTree with its
// std::list / heap-address tie-break as glibc really lays it out, IC_Angle, steered BRIEF with glibc's cosf / sinf --
// run as compiled code and pin the restatement in oracle_orb.cpp and the CUDA path.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../oracle.h"

namespace cv {

typedef unsigned char uchar;
enum { CV_8U = 0, CV_8UC1 = 0 };
enum { BORDER_CONSTANT = 0, BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16 };
enum { INTER_LINEAR = 1 };

inline int cvRound(double v) { return orc::cv_round(v); }
inline int cvRound(float v) { return orc::cv_round(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { return (int)std::floor(v); }
inline int cvCeil(double v) { return (int)std::ceil(v); }
#ifndef CV_PI
#define CV_PI 3.1415926535897932384626433832795
#endif
inline float fastAtan2(float y, float x) { return orc::fast_atan2(y, x); }

template <typename T>
struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    Point_& operator*=(float s) {  // Point_<float> *= float, ORBextractor.cpp:1050
        x = (T)(x * s);
        y = (T)(y * s);
        return *this;
    }
};
typedef Point_<float> Point2f;
typedef Point_<int> Point2i;
typedef Point2i Point;

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};
struct Rect {
    int x, y, width, height;
    Rect() : x(0), y(0), width(0), height(0) {}
    Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};

struct KeyPoint {  // same 28-byte layout as OpenCV
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
        : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
};

struct ZerosExpr {  // Mat::zeros(): assigned IN PLACE when the destination already has the size / type (OpenCV's MatExpr)
    int rows, cols, type;
};

class Mat {
public:
    int rows = 0, cols = 0;
    size_t step = 0;
    uchar* data = nullptr;

    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(Size sz, int type) { create(sz.height, sz.width, type); }
    Mat(int r, int c, int type, void* external, size_t step_) : rows(r), cols(c), step(step_), data((uchar*)external) { (void)type; }

    void create(int r, int c, int type) {
        (void)type;
        if (data && r == rows && c == cols) return;  // like cv::Mat::create: keeps the buffer (and an ROI view) if it fits
        rows = r;
        cols = c;
        step = (size_t)c;
        owner_ = std::make_shared<std::vector<uchar>>((size_t)r * c);
        data = owner_->data();
    }
    void release() {
        owner_.reset();
        data = nullptr;
        rows = cols = 0;
        step = 0;
    }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return CV_8UC1; }
    size_t step1() const { return step; }  // elements per row (8-bit single channel)
    Mat operator()(const Rect& r) const {
        Mat m(*this);
        m.data = data + (size_t)r.y * step + r.x;
        m.rows = r.height;
        m.cols = r.width;
        return m;
    }
    Mat rowRange(int a, int b) const { return (*this)(Rect(0, a, cols, b - a)); }
    Mat colRange(int a, int b) const { return (*this)(Rect(a, 0, b - a, rows)); }
    Mat clone() const {
        Mat m;
        if (empty()) return m;
        m.create(rows, cols, CV_8UC1);
        for (int r = 0; r < rows; ++r) std::memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)cols);
        return m;
    }
    static ZerosExpr zeros(int r, int c, int type) { return ZerosExpr{r, c, type}; }
    Mat& operator=(const ZerosExpr& z) {
        create(z.rows, z.cols, z.type);
        for (int r = 0; r < rows; ++r) std::memset(data + (size_t)r * step, 0, (size_t)cols);
        return *this;
    }
    uchar* ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar* ptr(int r = 0) const { return data + (size_t)r * step; }
    template <typename T>
    T& at(int r, int c) { return reinterpret_cast<T*>(data + (size_t)r * step)[c]; }
    template <typename T>
    const T& at(int r, int c) const { return reinterpret_cast<const T*>(data + (size_t)r * step)[c]; }

private:
    std::shared_ptr<std::vector<uchar>> owner_;
};

class _InputArray {
public:
    _InputArray(const Mat& m) : m_(&m) {}
    bool empty() const { return m_->empty(); }
    Mat getMat() const { return *m_; }

private:
    const Mat* m_;
};
class _OutputArray {
public:
    _OutputArray(Mat& m) : m_(&m) {}
    void release() const { m_->release(); }
    void create(int r, int c, int type) const { m_->create(r, c, type); }
    Mat getMat() const { return *m_; }

private:
    Mat* m_;
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

inline orc::Image view_of(const Mat& m) { return orc::Image{m.data, m.cols, m.rows, m.step}; }

// cv::resize(src, dst, dsize, 0, 0, INTER_LINEAR): dst already has dsize here (a view into the bordered level)
inline void resize(InputArray src, OutputArray dst, Size dsize, double, double, int interpolation) {
    assert(interpolation == INTER_LINEAR);
    (void)interpolation;
    const Mat s = src.getMat();
    dst.create(dsize.height, dsize.width, CV_8UC1);
    Mat d = dst.getMat();
    orc::resize_linear_u8(view_of(s), d.data, dsize.width, dsize.height, d.step);
}

inline int reflect101(int i, int n) {
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}
// cv::copyMakeBorder(src, dst, t, b, l, r, BORDER_REFLECT_101 [+ BORDER_ISOLATED]); src may be the interior view of dst
inline void copyMakeBorder(InputArray src_, OutputArray dst_, int top, int bottom, int left, int right, int borderType) {
    assert((borderType & ~BORDER_ISOLATED) == BORDER_REFLECT_101);
    (void)borderType;
    const Mat s = src_.getMat();
    dst_.create(s.rows + top + bottom, s.cols + left + right, CV_8UC1);
    Mat d = dst_.getMat();
    const bool aliased = s.data == d.data + (size_t)top * d.step + left;
    for (int y = 0; y < d.rows; ++y) {
        const int sy = reflect101(y - top, s.rows);
        for (int x = 0; x < d.cols; ++x) {
            const int sx = reflect101(x - left, s.cols);
            if (aliased && y - top == sy && x - left == sx) continue;  // the interior is already in place
            d.data[(size_t)y * d.step + x] = s.data[(size_t)sy * s.step + sx];
        }
    }
}

// cv::FAST(image, keypoints, threshold, nonmaxSuppression) -- TYPE_9_16
inline void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true) {
    const Mat m = image.getMat();
    std::vector<orc::KeyPoint> k;
    orc::fast9_detect(view_of(m), threshold, nonmaxSuppression, k);
    keypoints.clear();
    keypoints.reserve(k.size());
    for (const orc::KeyPoint& p : k) keypoints.push_back(KeyPoint(p.x, p.y, p.size, p.angle, p.response, p.octave, p.class_id));
}

// cv::GaussianBlur(src, dst, Size(7,7), 2, 2, BORDER_REFLECT_101), in place in the reference
inline void GaussianBlur(InputArray src, OutputArray dst, Size ksize, double sx, double sy, int borderType) {
    assert(ksize.width == 7 && ksize.height == 7 && sx == 2 && sy == 2 && borderType == BORDER_REFLECT_101);
    (void)ksize; (void)sx; (void)sy; (void)borderType;
    const Mat s = src.getMat();
    std::vector<uchar> tmp((size_t)s.rows * s.cols);
    orc::gaussian_blur7_sigma2(view_of(s), tmp.data(), (size_t)s.cols);
    dst.create(s.rows, s.cols, CV_8UC1);
    Mat d = dst.getMat();
    for (int r = 0; r < s.rows; ++r) std::memcpy(d.data + (size_t)r * d.step, tmp.data() + (size_t)r * s.cols, (size_t)s.cols);
}

struct KeyPointsFilter {  // only reached from ComputeKeyPointsOld (dead code, ORBextractor.cpp:809-979); kept so the file compiles
    static void retainBest(std::vector<KeyPoint>& k, int n) {
        if (n >= 0 && (int)k.size() > n) {
            std::stable_sort(k.begin(), k.end(), [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
            k.resize((size_t)n);
        }
    }
};

// ---- the slice of features2d / video that tracking.cpp:534-660 uses (oracle/ref_tracking_capi.cpp) --------------------
struct DMatch {
    int queryIdx, trainIdx, imgIdx;
    float distance;
    DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(3.4028235e38f) {}
};

template <typename T>
using Ptr = std::shared_ptr<T>;

// DescriptorMatcher::create("BruteForce-Hamming")->match(query, train, matches): one DMatch per query, lowest train index on ties
class DescriptorMatcher {
public:
    static Ptr<DescriptorMatcher> create(const std::string& type) {
        assert(type == "BruteForce-Hamming");
        (void)type;
        return std::make_shared<DescriptorMatcher>();
    }
    void match(const Mat& q, const Mat& t, std::vector<DMatch>& matches) const {
        matches.clear();
        if (q.empty() || t.empty()) return;  // empty train set: no matches
        assert(q.cols == 32 && t.cols == 32 && q.step == 32 && t.step == 32);
        std::vector<int32_t> idx((size_t)q.rows), dist((size_t)q.rows);
        orc::match_hamming(q.data, q.rows, t.data, t.rows, idx.data(), dist.data());
        matches.resize((size_t)q.rows);
        for (int i = 0; i < q.rows; ++i) {
            matches[i].queryIdx = i;
            matches[i].trainIdx = idx[i];
            matches[i].imgIdx = 0;
            matches[i].distance = (float)dist[i];
        }
    }
};

struct TermCriteria {
    enum { COUNT = 1, MAX_ITER = COUNT, EPS = 2 };
    int type, maxCount;
    double epsilon;
    TermCriteria(int t, int c, double e) : type(t), maxCount(c), epsilon(e) {}
};

// cv::calcOpticalFlowPyrLK(prev, next, prevPts, nextPts, status, err, winSize, maxLevel, criteria, flags, minEigThreshold)
inline void calcOpticalFlowPyrLK(const Mat& prev, const Mat& next, const std::vector<Point2f>& prevPts, std::vector<Point2f>& nextPts,
                                 std::vector<uchar>& status, Mat& err, Size winSize, int maxLevel, TermCriteria criteria, int flags,
                                 double minEigThreshold) {
    assert(flags == 0 && winSize.width == winSize.height && criteria.type == (TermCriteria::COUNT + TermCriteria::EPS));
    (void)flags;
    (void)err;  // the reference never reads the error matrices
    const int n = (int)prevPts.size();
    nextPts.assign((size_t)n, Point2f());
    status.assign((size_t)n, 0);
    if (!n) return;
    static_assert(sizeof(Point2f) == 8, "Point2f layout");
    std::vector<float> e((size_t)n);
    orc::calc_optical_flow_pyr_lk(view_of(prev), view_of(next), reinterpret_cast<const float*>(prevPts.data()), n,
                                  reinterpret_cast<float*>(nextPts.data()), status.data(), e.data(), winSize.width, maxLevel,
                                  criteria.maxCount, criteria.epsilon, minEigThreshold);
}

}  // namespace cv
