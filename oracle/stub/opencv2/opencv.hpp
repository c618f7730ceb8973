// Minimal stand-in for <opencv2/opencv.hpp>, written for this repo (TEST INFRASTRUCTURE ONLY).
//
// Purpose: let the reference's own translation units (ORBExtractor.cc, Camera.cc and the searchByStereo line ranges
// of ORBMatcher.cc) compile UNMODIFIED, from where they lie under /root/reference, without a C++ OpenCV install
// (none exists in this image).  Only the symbols those files touch are provided.  The three
// image primitives (resize / GaussianBlur / FAST) and undistortPoints forward to the plain-C restatements in
// oracle/orb_oracle.c, which tests/ pin bit-exactly against cv2 4.13.0.
#pragma once

#include <emmintrin.h>

#include <algorithm>
#include <cassert>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <utility>
#include <vector>

typedef unsigned char uchar;
typedef unsigned short ushort;

#define CV_8U 0
#define CV_16U 2
#define CV_32F 5
#define CV_64F 6

static inline int cvRound(double v) { return _mm_cvtsd_si32(_mm_set_sd(v)); }
static inline int cvRound(float v) { return _mm_cvtss_si32(_mm_set_ss(v)); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) {
  int i = (int)v;
  return i - (i > v);
}
static inline int cvFloor(float v) {
  int i = (int)v;
  return i - (i > v);
}
static inline int cvFloor(int v) { return v; }
static inline int cvCeil(double v) {
  int i = (int)v;
  return i + (i < v);
}
static inline int cvCeil(float v) {
  int i = (int)v;
  return i + (i < v);
}
static inline int cvCeil(int v) { return v; }

namespace cv
{

template <typename T> struct Point_
{
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
};
typedef Point_<float> Point2f;
typedef Point_<int> Point;

struct Size
{
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
};

struct KeyPoint
{
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
  KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
  KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1)
      : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c)
  {
  }
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

struct DMatch
{
  int queryIdx, trainIdx, imgIdx;
  float distance;
  DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(std::numeric_limits<float>::max()) {}
  DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
};

enum
{
  INTER_LINEAR = 1
};
enum
{
  BORDER_REFLECT_101 = 4
};
enum
{
  NORM_L1 = 2
};

class Mat
{
public:
  int rows, cols;
  uchar *data;
  size_t step;

  Mat() : rows(0), cols(0), data(nullptr), step(0), type_(CV_8U) {}
  Mat(int r, int c, int type) : rows(0), cols(0), data(nullptr), step(0), type_(type) { create(r, c, type); }
  // header over caller-owned memory (no copy), like cv::Mat(rows, cols, type, void*, step)
  Mat(int r, int c, int type, void *ext, size_t step_ = 0) : rows(r), cols(c), data((uchar *)ext), step(step_), type_(type)
  {
    if (!step) step = (size_t)c * elemSize();
  }

  void create(int r, int c, int type)
  {
    type_ = type;
    rows = r;
    cols = c;
    step = (size_t)c * elemSize();
    store_ = std::shared_ptr<uchar>(new uchar[step * (size_t)r + 16], std::default_delete<uchar[]>());
    data = store_.get();
  }
  int type() const { return type_; }
  size_t elemSize() const { return type_ == CV_8U ? 1 : (type_ == CV_16U ? 2 : (type_ == CV_64F ? 8 : 4)); }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }

  template <typename T> T &at(int r, int c) { return *(T *)(data + (size_t)r * step + (size_t)c * sizeof(T)); }
  template <typename T> const T &at(int r, int c) const { return *(const T *)(data + (size_t)r * step + (size_t)c * sizeof(T)); }
  template <typename T> T &at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
  template <typename T> const T &at(int i) const { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
  template <typename T> T *ptr(int r = 0) { return (T *)(data + (size_t)r * step); }
  template <typename T> const T *ptr(int r = 0) const { return (const T *)(data + (size_t)r * step); }

  Mat rowRange(int a, int b) const
  {
    if (a < 0 || b > rows || a > b) throw std::out_of_range("cv::Mat::rowRange");
    Mat m(*this);
    m.data = data + (size_t)a * step;
    m.rows = b - a;
    return m;
  }
  Mat colRange(int a, int b) const
  {
    if (a < 0 || b > cols || a > b) throw std::out_of_range("cv::Mat::colRange");
    Mat m(*this);
    m.data = data + (size_t)a * elemSize();
    m.cols = b - a;
    return m;
  }
  void copyTo(Mat &dst) const
  {
    if (dst.data == data && dst.rows == rows && dst.cols == cols) return;
    Mat out(rows, cols, type_);
    for (int r = 0; r < rows; ++r) std::memcpy(out.data + (size_t)r * out.step, data + (size_t)r * step, (size_t)cols * elemSize());
    dst = out;
  }
  Mat clone() const
  {
    Mat m;
    copyTo(m);
    return m;
  }
  void convertTo(Mat &dst, int rtype, double alpha = 1, double beta = 0) const
  {
    if (rtype < 0) rtype = type_;
    assert(rtype == CV_32F);
    Mat out(rows, cols, CV_32F);
    const float a = (float)alpha, b = (float)beta;
    const bool scaled = !(alpha == 1 && beta == 0);
    for (int r = 0; r < rows; ++r)
      for (int c = 0; c < cols; ++c)
      {
        float v = type_ == CV_8U ? (float)at<uchar>(r, c) : (type_ == CV_16U ? (float)at<ushort>(r, c) : at<float>(r, c));
        out.at<float>(r, c) = scaled ? v * a + b : v;
      }
    dst = out;
  }
  // Mat /= scalar on CV_32F (the RGB-D ctor's depthImg /= dScale, src/Frame.cc:131): OpenCV evaluates it as
  // convertTo(m, -1, 1. / s), i.e. a float multiply by (float)(1. / s) (cvtScale32f)
  Mat &operator/=(double s)
  {
    assert(type_ == CV_32F);
    const float a = (float)(1.0 / s);
    for (int r = 0; r < rows; ++r)
      for (int c = 0; c < cols; ++c) at<float>(r, c) = at<float>(r, c) * a;
    return *this;
  }
  static Mat ones(int r, int c, int type)
  {
    assert(type == CV_32F);
    Mat m(r, c, type);
    for (int i = 0; i < r; ++i)
      for (int j = 0; j < c; ++j) m.at<float>(i, j) = 1.f;
    return m;
  }
  static Mat zeros(int r, int c, int type)
  {
    Mat m(r, c, type);
    std::memset(m.data, 0, m.step * (size_t)r);
    return m;
  }

private:
  int type_;
  std::shared_ptr<uchar> store_;
};

// CV_32F arithmetic used by ORBMatcher::SAD (real OpenCV goes through MatExpr; values are identical)
static inline Mat operator*(const Mat &a, double s)
{
  Mat out(a.rows, a.cols, CV_32F);
  for (int r = 0; r < a.rows; ++r)
    for (int c = 0; c < a.cols; ++c) out.at<float>(r, c) = (float)(a.at<float>(r, c) * s);
  return out;
}
static inline Mat operator-(const Mat &a, const Mat &b)
{
  Mat out(a.rows, a.cols, CV_32F);
  for (int r = 0; r < a.rows; ++r)
    for (int c = 0; c < a.cols; ++c) out.at<float>(r, c) = a.at<float>(r, c) - b.at<float>(r, c);
  return out;
}
static inline double norm(const Mat &a, const Mat &b, int /*NORM_L1*/)
{
  double acc = 0;
  for (int r = 0; r < a.rows; ++r)
    for (int c = 0; c < a.cols; ++c) acc += std::fabs((double)(a.at<float>(r, c) - b.at<float>(r, c)));
  return acc;
}

struct NoArray
{
};
static inline NoArray noArray() { return NoArray(); }

// Implemented in oracle/stub/opencv_stub.cpp on top of oracle/orb_oracle.c
void resize(const Mat &src, Mat &dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void GaussianBlur(const Mat &src, Mat &dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_REFLECT_101);
void FAST(const Mat &image, std::vector<KeyPoint> &keypoints, int threshold, bool nonmaxSuppression = true);
void undistortPoints(const std::vector<Point2f> &src, std::vector<Point2f> &dst, const Mat &K, const Mat &dist, NoArray R, const Mat &P);

} // namespace cv
