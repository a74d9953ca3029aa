// tests/opencv_api_stub/opencv2/core/core.hpp -- TEST INFRASTRUCTURE, declarations only.
//
// This image has no OpenCV C++ headers, so "does the drop-in host code compile against the real OpenCV API?" (INTEGRATION.md route
// A, -DSVO_USE_REAL_OPENCV) is checked against this hand-written description of the OpenCV 3 / 4 API surface the host sources
// touch: the proxy classes cv::_InputArray / cv::_OutputArray (NOT Mat references, which is what the in-repo shim uses), cv::MatStep,
// cv::Mat, cv::KeyPoint, cv::DMatch, cv::Point_, the CV_* type macros.  Nothing is defined: tests/test_abi_cpu.py runs
// `g++ -fsyntax-only`, which type-checks every call the host sources make.  It is not a substitute for OpenCV.
#pragma once
#include <cstddef>
#include <vector>

#define CV_8U 0
#define CV_32F 5
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_64F 6
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)

namespace cv {

typedef unsigned char uchar;

template <typename _Tp>
class Point_ {
public:
    Point_();
    Point_(_Tp _x, _Tp _y);
    _Tp x, y;
};
template <typename _Tp>
Point_<_Tp>& operator*=(Point_<_Tp>& a, float b);
typedef Point_<int> Point2i;
typedef Point_<float> Point2f;
typedef Point2i Point;

template <typename _Tp, int cn>
class Vec {
public:
    Vec();
    Vec(_Tp v0, _Tp v1, _Tp v2);
    _Tp& operator[](int i);
    const _Tp& operator[](int i) const;
    _Tp val[cn];
};
typedef Vec<float, 3> Vec3f;

template <typename _Tp>
class Size_ {
public:
    Size_();
    Size_(_Tp _width, _Tp _height);
    _Tp width, height;
};
typedef Size_<int> Size;

struct MatStep {
    MatStep();
    explicit MatStep(size_t s);
    operator size_t() const;
    MatStep& operator=(size_t s);
    size_t* p;
    size_t buf[2];
};

class _InputArray;
class _OutputArray;

class Mat {
public:
    Mat();
    Mat(int rows, int cols, int type);
    Mat(int rows, int cols, int type, void* data, size_t step = 0);
    Mat(const Mat& m);
    ~Mat();
    Mat& operator=(const Mat& m);
    Mat clone() const;
    void create(int rows, int cols, int type);
    void release();
    bool empty() const;
    int type() const;
    size_t elemSize() const;
    bool isContinuous() const;
    uchar* ptr(int i0 = 0);
    const uchar* ptr(int i0 = 0) const;
    template <typename _Tp> _Tp* ptr(int i0 = 0);
    template <typename _Tp> const _Tp* ptr(int i0 = 0) const;
    template <typename _Tp> _Tp& at(int i0, int i1);
    template <typename _Tp> const _Tp& at(int i0, int i1) const;
    int flags, dims, rows, cols;
    uchar* data;
    MatStep step;
};

class _InputArray {
public:
    _InputArray();
    _InputArray(const Mat& m);
    Mat getMat(int idx = -1) const;
    bool empty() const;
    int type(int i = -1) const;
    int cols(int i = -1) const;
    int rows(int i = -1) const;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray();
    _OutputArray(Mat& m);
    void create(int rows, int cols, int type, int i = -1, bool allowTransposed = false, int fixedDepthMask = 0) const;
    void release() const;
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

class KeyPoint {
public:
    KeyPoint();
    KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0, int _class_id = -1);
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
};

class DMatch {
public:
    DMatch();
    DMatch(int _queryIdx, int _trainIdx, float _distance);
    int queryIdx, trainIdx, imgIdx;
    float distance;
};

}  // namespace cv
