// Stand-in for ORB_SLAM2_ROS2::Frame (TEST INFRASTRUCTURE ONLY): same member names / getters as the reference's
// Frame (include/ORB_SLAM2/Frame.h:263-274,340-347) restricted to what ORBMatcher::searchByStereo touches.
#pragma once
#include "ORB_SLAM2/ORBExtractor.h"

namespace ORB_SLAM2_ROS2
{
class VirtualFrame
{
};
class Frame : public VirtualFrame
{
public:
  typedef std::shared_ptr<Frame> SharedPtr;
  std::vector<cv::KeyPoint> mvFeatsLeft, mvFeatsRight;
  std::vector<cv::Mat> mvLeftDescriptor, mRightDescriptor;
  std::vector<double> mvDepths, mvFeatsRightU;
  cv::Mat mLeftIm, mRightIm;
  ORBExtractor::SharedPtr mpExtractorLeft, mpExtractorRight;
  int mnN = 0;

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
