// C-callable harness around the reference's OWN code (TEST INFRASTRUCTURE ONLY).
//
// Links against /root/reference/src/ORB_SLAM2/src/{ORBExtractor,Camera}.cc compiled unmodified and the searchByStereo
// ranges of ORBMatcher.cc (ref_stereo_tu.cpp), all built against oracle/stub.  Drives them exactly like the reference's
// Frame does (src/Frame.cc:85-111 and include/ORB_SLAM2/Frame.h:313-322): two extractor constructors in sequence, two
// std::threads running extract(), Camera::undistortPoints on the left keypoints, then ORBMatcher::searchByStereo.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <limits>
#include <memory>
#include <sstream>
#include <thread>
#include <vector>

#include <opencv2/opencv.hpp>
#include <rclcpp/rclcpp.hpp>

// The reference keeps scale factors / level count / template as process-wide private statics that are initialised once
// (src/ORBExtractor.cc:511-524).  To test several (nLevels, scaleFactor) configurations in one process the harness
// needs to clear them; the access-specifier override is applied to this harness TU only.
#define private public
#include "ORB_SLAM2/ORBExtractor.h"
#undef private
#include "ORB_SLAM2/Camera.h"
#include "ORB_SLAM2/Error.h"
#include "ORB_SLAM2/ORBMatcher.h"

#include "orb_oracle.h"
#include "ref_frame_standin.h"

using namespace ORB_SLAM2_ROS2;

namespace ref_private // ref_stereo_tu.cpp: wrappers around ORBMatcher's private statics
{
std::pair<std::size_t, int> best_match(const cv::Mat &desc, const std::vector<cv::Mat> &all, const std::vector<std::size_t> &cand, float &ratio);
void verify_angle(std::vector<cv::DMatch> &m, const std::vector<cv::KeyPoint> &k1, const std::vector<cv::KeyPoint> &k2);
} // namespace ref_private

namespace
{

cv::Mat wrap_u8(const uint8_t *p, int w, int h, size_t stride) { return cv::Mat(h, w, CV_8U, (void *)p, stride); }

void export_kps(const std::vector<cv::KeyPoint> &kps, const std::vector<cv::Mat> &descs, oracle_keypoint *okps, uint8_t *odesc, int cap)
{
  for (size_t i = 0; i < kps.size() && (int)i < cap; ++i)
  {
    std::memcpy(&okps[i], &kps[i], sizeof(oracle_keypoint));
    std::memcpy(odesc + 32 * i, descs[i].data, 32);
  }
}

struct StereoOut
{
  std::shared_ptr<Frame> frame;
  int nMatches = 0;
};

// Frame::Frame(stereo) + Frame::createStereo, minus the map/grid/BoW members that are outside the hot path
StereoOut run_stereo(const cv::Mat &l, const cv::Mat &r, int nFeatures, int nLevels, float scale, const std::string &tmpl, int iniTh, int minTh,
                     bool two_threads)
{
  StereoOut out;
  auto f = std::make_shared<Frame>();
  f->mLeftIm = l;
  f->mRightIm = r;
  f->mpExtractorLeft = std::make_shared<ORBExtractor>(f->mLeftIm, nFeatures, nLevels, scale, tmpl, iniTh, minTh);
  f->mpExtractorRight = std::make_shared<ORBExtractor>(f->mRightIm, nFeatures, nLevels, scale, tmpl, iniTh, minTh);
  if (two_threads)
  {
    std::thread lt(std::bind(&ORBExtractor::extract, f->mpExtractorLeft.get(), std::ref(f->mvFeatsLeft), std::ref(f->mvLeftDescriptor)));
    std::thread rt(std::bind(&ORBExtractor::extract, f->mpExtractorRight.get(), std::ref(f->mvFeatsRight), std::ref(f->mRightDescriptor)));
    lt.join();
    rt.join();
  }
  else
  {
    f->mpExtractorLeft->extract(f->mvFeatsLeft, f->mvLeftDescriptor);
    f->mpExtractorRight->extract(f->mvFeatsRight, f->mRightDescriptor);
  }
  Camera::undistortPoints(f->mvFeatsLeft);
  ORBMatcher matcher;
  out.nMatches = matcher.searchByStereo(f);
  f->mnN = out.nMatches;
  out.frame = f;
  return out;
}

} // namespace

extern "C"
{

  // forget the process-wide statics so that a different (nLevels, scaleFactor, template) can be configured
  void ref_reset()
  {
    ORBExtractor::mbScaleInit = false;
    ORBExtractor::mvfScaledFactors.clear();
    ORBExtractor::mvnFeatures.clear();
    ORBExtractor::mbTemInit = false;
    ORBExtractor::mvBriefTem.clear();
    ORBExtractor::mbMaxColInit = false;
    ORBExtractor::mvMaxColIdx.clear();
  }

  // System::setSetting (src/System.cc:27-73): Camera statics, K, and the 4- or 5-element distortion vector
  void ref_set_camera(float fx, float fy, float cx, float cy, float bl, const float *dist5)
  {
    Camera::mfFx = fx;
    Camera::mfFy = fy;
    Camera::mfCx = cx;
    Camera::mfCy = cy;
    Camera::mfBl = bl;
    Camera::mfBf = Camera::mfFx * Camera::mfBl;
    cv::Mat K = cv::Mat::zeros(3, 3, CV_32F);
    K.at<float>(0, 0) = fx;
    K.at<float>(0, 2) = cx;
    K.at<float>(1, 1) = fy;
    K.at<float>(1, 2) = cy;
    K.at<float>(2, 2) = 1.f;
    Camera::mK = K;
    int nd = (dist5 && dist5[4] != 0) ? 5 : 4;
    cv::Mat D = cv::Mat::zeros(nd, 1, CV_32F);
    for (int i = 0; i < nd; ++i) D.at<float>(i) = dist5 ? dist5[i] : 0.f;
    Camera::mDistCoeff = D;
  }

  float ref_get_bf() { return Camera::mfBf; }

  // ORBExtractor ctor + extract.  pyr_out (optional) receives the pyramid levels densely packed one after another;
  // lw/lh (optional, [nLevels]) their sizes; sf_out the static scale factors.  Returns n, or <0:
  // -1 ImageSizeError, -2 FileNotOpenError, -3 any other exception.
  int ref_extract(const uint8_t *img, int w, int h, size_t stride, int nFeatures, int nLevels, float scale, const char *tmpl, int iniTh, int minTh,
                  oracle_keypoint *kps, uint8_t *desc, int cap, uint8_t *pyr_out, int *lw, int *lh, float *sf_out)
  {
    try
    {
      cv::Mat im = wrap_u8(img, w, h, stride);
      ORBExtractor ex(im, nFeatures, nLevels, scale, tmpl, iniTh, minTh);
      std::vector<cv::KeyPoint> k;
      std::vector<cv::Mat> d;
      ex.extract(k, d);
      export_kps(k, d, kps, desc, cap);
      const auto &pyr = ex.getPyramid();
      size_t off = 0;
      for (size_t l = 0; l < pyr.size(); ++l)
      {
        if (lw) lw[l] = pyr[l].cols;
        if (lh) lh[l] = pyr[l].rows;
        if (sf_out) sf_out[l] = ORBExtractor::getScaledFactors()[l];
        if (pyr_out)
          for (int y = 0; y < pyr[l].rows; ++y, off += (size_t)pyr[l].cols) std::memcpy(pyr_out + off, pyr[l].data + (size_t)y * pyr[l].step, (size_t)pyr[l].cols);
      }
      return (int)k.size();
    }
    catch (const ImageSizeError &)
    {
      return -1;
    }
    catch (const FileNotOpenError &)
    {
      return -2;
    }
    catch (const std::exception &e)
    {
      std::fprintf(stderr, "ref_extract: %s\n", e.what());
      return -3;
    }
  }

  // the blurred level images (private mvBriefMat), densely packed -- only used to cross-check the blur restatement
  int ref_blurred_pyramid(const uint8_t *img, int w, int h, size_t stride, int nFeatures, int nLevels, float scale, const char *tmpl, uint8_t *out)
  {
    try
    {
      cv::Mat im = wrap_u8(img, w, h, stride);
      ORBExtractor ex(im, nFeatures, nLevels, scale, tmpl, 20, 7);
      size_t off = 0;
      for (auto &m : ex.mvBriefMat)
        for (int y = 0; y < m.rows; ++y, off += (size_t)m.cols) std::memcpy(out + off, m.data + (size_t)y * m.step, (size_t)m.cols);
      return 0;
    }
    catch (const std::exception &)
    {
      return -3;
    }
  }

  // the reference Quadtree on its own (src/ORBExtractor.cc:126-192)
  int ref_quadtree(int roi_w, int roi_h, int n, const float *xs, const float *ys, const float *resp, int need, int *out_idx)
  {
    std::vector<cv::KeyPoint> kps(n);
    for (int i = 0; i < n; ++i) kps[i] = cv::KeyPoint(xs[i], ys[i], 7.f, -1, resp[i]);
    Quadtree qt(roi_w, roi_h, kps, (unsigned)need);
    qt.split();
    int k = 0;
    for (auto id : qt.getFeatIdxs()) out_idx[k++] = (int)id;
    return k;
  }

  // Frame::createStereo.  Returns nMatches (>=0) or a negative error as in ref_extract.
  int ref_stereo(const uint8_t *left, const uint8_t *right, int w, int h, size_t stride, int nFeatures, int nLevels, float scale, const char *tmpl, int iniTh,
                 int minTh, oracle_keypoint *kl, uint8_t *dl, int *nl, oracle_keypoint *kr, uint8_t *dr, int *nr, double *u_right, double *depth, int cap)
  {
    try
    {
      StereoOut o = run_stereo(wrap_u8(left, w, h, stride), wrap_u8(right, w, h, stride), nFeatures, nLevels, scale, tmpl, iniTh, minTh, true);
      auto &f = *o.frame;
      export_kps(f.mvFeatsLeft, f.mvLeftDescriptor, kl, dl, cap);
      export_kps(f.mvFeatsRight, f.mRightDescriptor, kr, dr, cap);
      *nl = (int)f.mvFeatsLeft.size();
      *nr = (int)f.mvFeatsRight.size();
      for (int i = 0; i < *nl && i < cap; ++i)
      {
        u_right[i] = f.mvFeatsRightU[i];
        depth[i] = f.mvDepths[i];
      }
      return o.nMatches;
    }
    catch (const ImageSizeError &)
    {
      return -1;
    }
    catch (const FileNotOpenError &)
    {
      return -2;
    }
    catch (const std::exception &e)
    {
      std::fprintf(stderr, "ref_stereo: %s\n", e.what());
      return -3;
    }
  }

  // Frame::Frame(colorImg, depthImg, ...) (src/Frame.cc:125-159): the reference's own ctor body on a gray image and a raw depth image
  // (uint16 or float32).  kps = mvFeatsLeft (undistorted), depth / u_right = mvDepths / mvFeatsRightU.  Returns the keypoint count.
  int ref_rgbd(const uint8_t *gray, int w, int h, size_t stride, const void *depth, int depth_is_float, size_t depth_stride_bytes, float dScale, int nFeatures,
               int nLevels, float scale, const char *tmpl, int iniTh, int minTh, oracle_keypoint *kps, uint8_t *desc, double *u_right, double *depth_out, int cap)
  {
    try
    {
      Frame f;
      f.mfMinU = 0, f.mfMinV = 0, f.mfMaxU = (float)w, f.mfMaxV = (float)h; // VirtualFrame ctor (Frame.h:33-43) without distortion; the grid is not exported here
      if (Camera::mDistCoeff.at<float>(0) != 0.f)
      {
        std::vector<cv::KeyPoint> c(2);
        c[0].pt = cv::Point2f(0.f, 0.f);
        c[1].pt = cv::Point2f((float)w, (float)h);
        Camera::undistortPoints(c);
        f.mfMinU = c[0].pt.x, f.mfMinV = c[0].pt.y, f.mfMaxU = c[1].pt.x, f.mfMaxV = c[1].pt.y;
      }
      cv::Mat g = wrap_u8(gray, w, h, stride);
      cv::Mat d(h, w, depth_is_float ? CV_32F : CV_16U, (void *)depth, depth_stride_bytes);
      f.rgbdCtorBody(g, d, nFeatures, tmpl, iniTh, minTh, dScale, nLevels, scale);
      export_kps(f.mvFeatsLeft, f.mvLeftDescriptor, kps, desc, cap);
      const int n = (int)f.mvFeatsLeft.size();
      for (int i = 0; i < n && i < cap; ++i)
      {
        u_right[i] = f.mvFeatsRightU[i];
        depth_out[i] = f.mvDepths[i];
      }
      return n;
    }
    catch (const ImageSizeError &)
    {
      return -1;
    }
    catch (const FileNotOpenError &)
    {
      return -2;
    }
    catch (const std::exception &e)
    {
      std::fprintf(stderr, "ref_rgbd: %s\n", e.what());
      return -3;
    }
  }

  // Camera::undistortPoints (src/Camera.cc:29-39) on raw coordinates
  void ref_undistort(float *xy, int n)
  {
    std::vector<cv::KeyPoint> kps(n);
    for (int i = 0; i < n; ++i) kps[i].pt = cv::Point2f(xy[2 * i], xy[2 * i + 1]);
    Camera::undistortPoints(kps);
    for (int i = 0; i < n; ++i)
    {
      xy[2 * i] = kps[i].pt.x;
      xy[2 * i + 1] = kps[i].pt.y;
    }
  }

  // VirtualFrame::initGrid (src/Frame.cc:53-69) on the given keypoints -> CSR; returns rows * cols, or -1 if cap is small
  int ref_init_grid(const oracle_keypoint *kps, int n, float min_u, float min_v, float max_u, float max_v, int *rows, int *cols, int *start, int cap_cells,
                    int *entries)
  {
    VirtualFrame f;
    f.mvFeatsLeft.resize(n);
    for (int i = 0; i < n; ++i) std::memcpy(&f.mvFeatsLeft[i], &kps[i], sizeof(oracle_keypoint));
    f.mfMinU = min_u, f.mfMinV = min_v, f.mfMaxU = max_u, f.mfMaxV = max_v;
    f.initGrid();
    *rows = (int)f.mGrids.size();
    *cols = *rows ? (int)f.mGrids[0].size() : 0;
    if (*rows * *cols > cap_cells) return -1;
    int k = 0, c = 0;
    for (auto &row : f.mGrids)
      for (auto &cell : row)
      {
        start[c++] = k;
        for (auto id : cell) entries[k++] = (int)id;
      }
    start[c] = k;
    return c;
  }

  // The inner step of both ORBMatcher::searchByProjection overloads, composed from the reference's own compiled
  // functions: VirtualFrame::initGrid + findFeaturesInArea (src/Frame.cc:53-69,286-311), the exclusion filter of
  // src/ORBMatcher.cc:322-331 (a plain copy_if here) and ORBMatcher::getBestMatch (:967-990).
  void ref_search_in_area(const oracle_keypoint *kps, const uint8_t *desc, int n_kps, float min_u, float min_v, float max_u, float max_v, const float *sf,
                          int n_levels, const oracle_area_query *q, const uint8_t *q_desc, int n_q, const uint8_t *exclude, int *best_idx, int *best_dist,
                          float *ratio, int *n_cand)
  {
    VirtualFrame f;
    f.mvFeatsLeft.resize(n_kps);
    std::vector<cv::Mat> descs(n_kps);
    for (int i = 0; i < n_kps; ++i)
    {
      std::memcpy(&f.mvFeatsLeft[i], &kps[i], sizeof(oracle_keypoint));
      descs[i] = cv::Mat(1, 32, CV_8U, (void *)(desc + 32 * (size_t)i), 32);
    }
    f.mfMinU = min_u, f.mfMinV = min_v, f.mfMaxU = max_u, f.mfMaxV = max_v;
    VirtualFrame::mvfScaledFactors.assign(sf, sf + n_levels);
    f.initGrid();
    // findFeaturesInArea reads cell floor(mfMaxU / 64) / row floor(mfMaxV / 48), which lies one past the grid whenever the
    // bound is a multiple of the cell size (320x240, 640x480 ...): out-of-range vector reads in the reference.  One empty
    // guard row and column turn them into reads of empty cells (= the clipping the oracle and the CUDA path apply).
    for (auto &row : f.mGrids) row.emplace_back();
    f.mGrids.emplace_back(f.mGrids.empty() ? 1 : f.mGrids[0].size());
    for (int i = 0; i < n_q; ++i)
    {
      cv::KeyPoint kp;
      kp.pt = cv::Point2f(q[i].x, q[i].y);
      kp.octave = q[i].octave;
      std::vector<std::size_t> cand = f.findFeaturesInArea(kp, q[i].radius, q[i].min_level, q[i].max_level), kept;
      for (auto id : cand)
        if (!exclude || !exclude[id]) kept.push_back(id);
      n_cand[i] = (int)kept.size();
      best_idx[i] = -1;
      best_dist[i] = std::numeric_limits<int>::max();
      ratio[i] = 0.f;
      if (kept.empty()) continue;
      cv::Mat d(1, 32, CV_8U, (void *)(q_desc + 32 * (size_t)i), 32);
      float r = 0.f;
      auto best = ref_private::best_match(d, descs, kept, r);
      best_idx[i] = (int)best.first;
      best_dist[i] = best.second;
      ratio[i] = r;
    }
  }

  // ORBMatcher::verifyAngle (src/ORBMatcher.cc:1013-1051), in place; returns the new count
  int ref_verify_angle(int n, int *query_idx, int *train_idx, float *distance, const oracle_keypoint *kps1, int n1, const oracle_keypoint *kps2, int n2)
  {
    std::vector<cv::KeyPoint> k1(n1), k2(n2);
    for (int i = 0; i < n1; ++i) std::memcpy(&k1[i], &kps1[i], sizeof(oracle_keypoint));
    for (int i = 0; i < n2; ++i) std::memcpy(&k2[i], &kps2[i], sizeof(oracle_keypoint));
    std::vector<cv::DMatch> m(n);
    for (int i = 0; i < n; ++i)
    {
      m[i].queryIdx = query_idx[i];
      m[i].trainIdx = train_idx[i];
      m[i].distance = distance[i];
    }
    ref_private::verify_angle(m, k1, k2);
    for (size_t i = 0; i < m.size(); ++i)
    {
      query_idx[i] = m[i].queryIdx;
      train_idx[i] = m[i].trainIdx;
      distance[i] = m[i].distance;
    }
    return (int)m.size();
  }

  // CPU baseline: process `n_frames` stereo frames drawn round-robin from a pool of `pool` pairs (each w x h, dense),
  // with `workers` frames in flight; every frame uses the reference's own two extractor threads (Frame.cc:100-105),
  // so up to 2*workers host threads are busy.  Returns wall seconds; *matches_out accumulates nMatches as a checksum.
  double ref_bench_stereo(const uint8_t *left_pool, const uint8_t *right_pool, int pool, int w, int h, int nFeatures, int nLevels, float scale,
                          const char *tmpl, int iniTh, int minTh, int n_frames, int workers, long *matches_out)
  {
    const size_t fsz = (size_t)w * (size_t)h;
    // one warm frame initialises the process-wide statics before any concurrency
    run_stereo(wrap_u8(left_pool, w, h, (size_t)w), wrap_u8(right_pool, w, h, (size_t)w), nFeatures, nLevels, scale, tmpl, iniTh, minTh, true);
    std::atomic<int> next(0);
    std::atomic<long> matches(0);
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool_threads;
    for (int t = 0; t < workers; ++t)
      pool_threads.emplace_back(
          [&]()
          {
            for (;;)
            {
              int i = next.fetch_add(1);
              if (i >= n_frames) break;
              const uint8_t *l = left_pool + fsz * (size_t)(i % pool), *r = right_pool + fsz * (size_t)(i % pool);
              StereoOut o = run_stereo(wrap_u8(l, w, h, (size_t)w), wrap_u8(r, w, h, (size_t)w), nFeatures, nLevels, scale, tmpl, iniTh, minTh, true);
              matches += o.nMatches;
            }
          });
    for (auto &t : pool_threads) t.join();
    auto t1 = std::chrono::steady_clock::now();
    if (matches_out) *matches_out = matches.load();
    return std::chrono::duration<double>(t1 - t0).count();
  }

} // extern "C"
