// Translation unit that compiles the reference's stereo matcher VERBATIM (TEST INFRASTRUCTURE ONLY).
//
// ORBMatcher.cc as a whole needs Frame/KeyFrame/MapPoint/DBoW3/g2o; its stereo functions do not.  The Makefile
// extracts lines 18-81 (searchByStereo), 841-1011 (pixelSADMatch, SAD, createRowIndexDB, descDistance, getBestMatch,
// getPitch) and 1086-1093 (static constants) of /root/reference/src/ORB_SLAM2/src/ORBMatcher.cc with sed into a
// temporary file (REF_STEREO_INC, never stored in this repo) that is #included below, after the reference's REAL
// ORBMatcher.h / Camera.h and a stand-in Frame exposing exactly the members those lines touch
// (include/ORB_SLAM2/Frame.h:263-274,340-347).
#include "ORB_SLAM2/Camera.h"
#include "ORB_SLAM2/ORBMatcher.h"

#include "ref_frame_standin.h"

namespace ORB_SLAM2_ROS2
{
#include REF_STEREO_INC
} // namespace ORB_SLAM2_ROS2
