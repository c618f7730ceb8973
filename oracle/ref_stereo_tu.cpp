// Translation unit that compiles the reference's matcher / frame-grid functions VERBATIM (TEST INFRASTRUCTURE ONLY).
//
// ORBMatcher.cc and Frame.cc as a whole need KeyFrame/MapPoint/DBoW3/g2o; the functions below do not.  The Makefile
// extracts with sed, into temporary files that are never stored in this repo:
//   ref_orbmatcher_ranges.inc  ORBMatcher.cc 18-81 (searchByStereo), 841-1051 (pixelSADMatch, SAD, createRowIndexDB,
//                              descDistance, getBestMatch, getPitch, verifyAngle), 1086-1093 (static constants)
//   ref_frame_cc_ranges.inc    Frame.cc 53-69 (VirtualFrame::initGrid), 286-311 (VirtualFrame::findFeaturesInArea),
//                              355-359 (static members)
//   ref_frame_rgbd_body.inc    Frame.cc 130-158 (the body of the RGB-D Frame ctor: convertTo, /= dScale, extractor, undistortion,
//                              initGrid, the depth lookup loop)
//   ref_frame_h_ranges.inc     Frame.h 204, 207 (getScaledFactor, getScaledFactor2; included by ref_frame_standin.h)
// They are #included below after the reference's REAL ORBMatcher.h / Camera.h and a stand-in VirtualFrame/Frame exposing
// exactly the members those lines touch.  getBestMatch / verifyAngle are private statics of ORBMatcher: the
// access-specifier override applies to this TU only and the C wrappers at the bottom expose them to the harness.
#include <limits>

#include "ORB_SLAM2/Camera.h"
#include "ORB_SLAM2/ORBExtractor.h"
#define private public
#include "ORB_SLAM2/ORBMatcher.h"
#undef private

#include "ref_frame_standin.h"

namespace ORB_SLAM2_ROS2
{
#include "ref_orbmatcher_ranges.inc"
#include "ref_frame_cc_ranges.inc"

// Frame::Frame(colorImg, depthImg, ...) (src/Frame.cc:125-159): the initialiser list (:127-128) sets mLeftIm and the image bounds
// (done by the harness); the body is the reference's own lines
void Frame::rgbdCtorBody(cv::Mat colorImg, cv::Mat depthImg, int nFeatures, const std::string &briefFp, int maxThresh, int minThresh, float dScale,
                         int nLevels, float scale)
{
  mLeftIm = colorImg;
#include "ref_frame_rgbd_body.inc"
}
} // namespace ORB_SLAM2_ROS2

namespace ref_private
{
std::pair<std::size_t, int> best_match(const cv::Mat &desc, const std::vector<cv::Mat> &all, const std::vector<std::size_t> &cand, float &ratio)
{
  return ORB_SLAM2_ROS2::ORBMatcher::getBestMatch(desc, all, cand, ratio);
}
void verify_angle(std::vector<cv::DMatch> &m, const std::vector<cv::KeyPoint> &k1, const std::vector<cv::KeyPoint> &k2)
{
  ORB_SLAM2_ROS2::ORBMatcher::verifyAngle(m, k1, k2);
}
} // namespace ref_private
