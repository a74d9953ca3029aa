// include/cvcompat/opencv2/core.hpp -- minimal stand-in for the OpenCV C++ types that cross the drop-in boundary.
//
// The reference's classes take cv::Mat / cv::KeyPoint / cv::Point2f / cv::InputArray (include/lzb_vio/ORBextractor.h:41-43,
// tracking.h:52-59).  This image has no OpenCV C++, so the host classes are compiled against this shim;
// with a real OpenCV on the include path, build with -DSVO_USE_REAL_OPENCV and this file is not used.  Only containers
// are provided -- no OpenCV algorithm is reimplemented here (those live in the CUDA kernels).
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

// the type codes are macros in OpenCV; same values here
#ifndef CV_8U
#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_32FC1 5
#define CV_64F 6
#define CV_64FC1 6
#define CV_32FC3 21
#endif

namespace cv {

typedef unsigned char uchar;

template <typename T>
struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    Point_& operator*=(T s) {
        x *= s;
        y *= s;
        return *this;
    }
};
typedef Point_<float> Point2f;
typedef Point_<int> Point2i;
typedef Point2i Point;

// cv::Vec3f (what Tracking::rotationMatrixToEulerAngles returns, tracking.h:99)
struct Vec3f {
    float val[3];
    Vec3f() : val{0, 0, 0} {}
    Vec3f(float a, float b, float c) : val{a, b, c} {}
    float& operator[](int i) { return val[i]; }
    const float& operator[](int i) const { return val[i]; }
};

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};

// 28 bytes, same field order as OpenCV's cv::KeyPoint
struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
        : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
};
static_assert(sizeof(KeyPoint) == 28, "KeyPoint must match cv::KeyPoint");

struct DMatch {
    int queryIdx, trainIdx, imgIdx;
    float distance;
    DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(3.4e38f) {}
    DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(0), distance(d) {}
};

// reference-counted 2-D matrix of 8-bit, 32-bit float (1 or 3 channels) or 64-bit float elements
class Mat {
public:
    int rows = 0, cols = 0;
    size_t step = 0;
    uchar* data = nullptr;

    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, void* external, size_t step_ = 0) : rows(r), cols(c), data((uchar*)external), type_(type) {
        step = step_ ? step_ : (size_t)c * elemSize();
    }
    void create(int r, int c, int type) {
        if (r == rows && c == cols && type == type_ && data && owner_) return;
        type_ = type;
        rows = r;
        cols = c;
        step = (size_t)c * elemSize();
        owner_ = std::shared_ptr<std::vector<uchar>>(new std::vector<uchar>((size_t)r * step));
        data = owner_->data();
    }
    void release() {
        owner_.reset();
        data = nullptr;
        rows = cols = 0;
        step = 0;
    }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    // cv::_InputArray::getMat() / cv::_OutputArray::getMat(): here the array types ARE Mat references, so this is the identity
    Mat getMat() const { return *this; }
    int type() const { return type_; }
    size_t elemSize() const { return type_ == CV_32F ? 4 : type_ == CV_64F ? 8 : type_ == CV_32FC3 ? 12 : 1; }
    bool isContinuous() const { return step == (size_t)cols * elemSize(); }
    Mat clone() const {
        Mat m;
        if (empty()) return m;
        m.create(rows, cols, type_);
        for (int r = 0; r < rows; ++r) std::memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)cols * elemSize());
        return m;
    }
    template <typename T>
    T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step); }
    template <typename T>
    const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }
    uchar* ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar* ptr(int r = 0) const { return data + (size_t)r * step; }
    template <typename T>
    T& at(int r, int c) { return ptr<T>(r)[c]; }
    template <typename T>
    const T& at(int r, int c) const { return ptr<T>(r)[c]; }

private:
    int type_ = CV_8UC1;
    std::shared_ptr<std::vector<uchar>> owner_;
};

typedef const Mat& InputArray;
typedef Mat& OutputArray;

}  // namespace cv
