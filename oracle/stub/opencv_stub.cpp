// Implementation of the image primitives declared by oracle/stub/opencv2/opencv.hpp (TEST INFRASTRUCTURE ONLY).
// They forward to the plain-C restatements in oracle/orb_oracle.c, which are pinned against cv2 4.13.0 by tests/.
#include <opencv2/opencv.hpp>

#include "../orb_oracle.h"

namespace cv
{

void resize(const Mat &src, Mat &dst, Size dsize, double, double, int interpolation)
{
  assert(interpolation == INTER_LINEAR && src.type() == CV_8U);
  (void)interpolation;
  Mat out(dsize.height, dsize.width, CV_8U);
  oracle_resize_linear_u8(src.data, src.cols, src.rows, src.step, out.data, dsize.width, dsize.height, out.step);
  dst = out;
}

void GaussianBlur(const Mat &src, Mat &dst, Size ksize, double sigmaX, double sigmaY, int borderType)
{
  assert(ksize.width == 7 && ksize.height == 7 && sigmaX == 2 && sigmaY == 2 && borderType == BORDER_REFLECT_101);
  (void)ksize;
  (void)sigmaX;
  (void)sigmaY;
  (void)borderType;
  Mat out(src.rows, src.cols, CV_8U);
  oracle_gaussian_blur7_u8(src.data, src.cols, src.rows, src.step, out.data, out.step);
  dst = out;
}

void FAST(const Mat &image, std::vector<KeyPoint> &keypoints, int threshold, bool nonmaxSuppression)
{
  assert(nonmaxSuppression);
  (void)nonmaxSuppression;
  keypoints.clear();
  const int cap = (image.rows / 2 + 1) * (image.cols / 2 + 1);
  std::vector<int> xs(cap), ys(cap), sc(cap);
  int n = oracle_fast9_nms(image.data, image.cols, image.rows, image.step, threshold, xs.data(), ys.data(), sc.data(), cap);
  for (int i = 0; i < n; ++i) keypoints.push_back(KeyPoint((float)xs[i], (float)ys[i], 7.f, -1, (float)sc[i]));
}

void undistortPoints(const std::vector<Point2f> &src, std::vector<Point2f> &dst, const Mat &K, const Mat &dist, NoArray, const Mat &P)
{
  (void)P; // the reference passes P == K (Camera.cc:36)
  std::vector<float> xy(2 * src.size());
  for (size_t i = 0; i < src.size(); ++i)
  {
    xy[2 * i] = src[i].x;
    xy[2 * i + 1] = src[i].y;
  }
  float d[5] = {0, 0, 0, 0, 0};
  int nd = dist.rows * dist.cols;
  for (int i = 0; i < nd && i < 5; ++i) d[i] = dist.at<float>(i);
  oracle_undistort_points(xy.data(), (int)src.size(), K.at<float>(0, 0), K.at<float>(1, 1), K.at<float>(0, 2), K.at<float>(1, 2), d, nd);
  dst.resize(src.size());
  for (size_t i = 0; i < dst.size(); ++i) dst[i] = Point2f(xy[2 * i], xy[2 * i + 1]);
}

} // namespace cv
