// Stand-in for ORB_SLAM2_ROS2::VirtualFrame / Frame (TEST INFRASTRUCTURE ONLY): same member names / getters as the
// reference's classes (include/ORB_SLAM2/Frame.h:28,204-207,263-299,340-347) restricted to what the compiled line ranges
// touch: ORBMatcher::searchByStereo, VirtualFrame::initGrid, VirtualFrame::findFeaturesInArea and the RGB-D Frame ctor.  The bodies of
// getScaledFactor / getScaledFactor2 are the reference's own lines (Frame.h:204,207), extracted with sed at build time
// into ref_frame_h_ranges.inc (never stored in this repo).
#pragma once
#include "ORB_SLAM2/ORBExtractor.h"

namespace ORB_SLAM2_ROS2
{
class VirtualFrame
{
public:
  typedef std::vector<std::vector<std::vector<std::size_t>>> GridsType;
  std::vector<cv::KeyPoint> mvFeatsLeft;
  float mfMaxU = 0, mfMaxV = 0, mfMinU = 0, mfMinV = 0;
  GridsType mGrids;
  static std::vector<float> mvfScaledFactors;
  static bool mbScaled;
  static std::size_t mnNextID;
  static unsigned mnGridHeight;
  static unsigned mnGridWidth;

  void initGrid();
  std::vector<std::size_t> findFeaturesInArea(const cv::KeyPoint &kp, float radius, int minNLevel, int maxNLevel);
#include "ref_frame_h_ranges.inc"
};
class Frame : public VirtualFrame
{
public:
  typedef std::shared_ptr<Frame> SharedPtr;
  std::vector<cv::KeyPoint> mvFeatsRight;
  std::vector<cv::Mat> mvLeftDescriptor, mRightDescriptor;
  std::vector<double> mvDepths, mvFeatsRightU;
  cv::Mat mLeftIm, mRightIm;
  ORBExtractor::SharedPtr mpExtractorLeft, mpExtractorRight;
  int mnN = 0;
  // members the RGB-D ctor body touches (include/ORB_SLAM2/Frame.h:281,31) + the body itself (src/Frame.cc:130-158, compiled
  // verbatim from ref_frame_rgbd_body.inc by ref_stereo_tu.cpp)
  std::vector<void *> mvpMapPoints;
  std::size_t mnID = 0;
  void rgbdCtorBody(cv::Mat colorImg, cv::Mat depthImg, int nFeatures, const std::string &briefFp, int maxThresh, int minThresh, float dScale, int nLevels,
                    float scale);

  const std::vector<cv::KeyPoint> &getLeftKeyPoints() const { return mvFeatsLeft; }
  const std::vector<cv::KeyPoint> &getRightKeyPoints() const { return mvFeatsRight; }
  const std::vector<cv::Mat> &getLeftDescriptor() const { return mvLeftDescriptor; }
  const std::vector<cv::Mat> &getRightDescriptor() const { return mRightDescriptor; }
  const cv::Mat &getLeftImage() const { return mLeftIm; }
  const cv::Mat &getRightImage() const { return mRightIm; }
  const std::vector<cv::Mat> &getLeftPyramid() const { return mpExtractorLeft->getPyramid(); }
  const std::vector<cv::Mat> &getRightPyramid() const { return mpExtractorRight->getPyramid(); }
};
} // namespace ORB_SLAM2_ROS2
